#include "train_walk.inl"
template int train_walk_t<double>(smore_model_s*, const smore_train_params*, int, int);
template int train_hpe_t<double>(smore_model_s*, const smore_train_params*);
