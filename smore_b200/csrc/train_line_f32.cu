#include "train_line.inl"
template int train_line_t<float>(smore_model_s*, const smore_train_params*);
