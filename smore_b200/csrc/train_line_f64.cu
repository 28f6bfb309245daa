#include "train_line.inl"
template int train_line_t<double>(smore_model_s*, const smore_train_params*);
template int train_line_exchange_t<double>(smore_model_s**, int, const smore_train_params*, ExchTransport&);
template int train_mf_t<double>(smore_model_s*, const smore_train_params*);
template int train_line_block_t<double>(smore_model_s*, const smore_train_params*, int, void*, uint64_t);
