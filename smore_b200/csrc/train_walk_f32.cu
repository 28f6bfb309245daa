#include "train_walk.inl"
template int train_walk_t<float>(smore_model_s*, const smore_train_params*, int, int);
template int train_hpe_t<float>(smore_model_s*, const smore_train_params*);
