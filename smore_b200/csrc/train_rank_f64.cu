#include "train_rank.inl"
template int train_ranking_t<double>(smore_model_s*, const smore_train_params*, int);
template int train_bpr_block_t<double>(smore_model_s*, const smore_train_params*, int, void*, uint64_t);
