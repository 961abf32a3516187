// instantiates the adapter against the mocks: a syntax/type check, nothing more
#include "ccsd_t_gpu.h"
template class mpqc::lcao::CCSD_T_GPU<TA::TensorD, TA::DensePolicy>;
int main() { mpqc::KeyVal kv; mpqc::lcao::CCSD_T_GPU<TA::TensorD, TA::SparsePolicy> w(kv); (void)w; return 0; }
