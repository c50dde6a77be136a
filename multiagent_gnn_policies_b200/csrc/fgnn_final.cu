// fgnn_final.cu -- compiled once per (FGNN_K, FGNN_HP) pair: -DFGNN_K=<1..4> -DFGNN_HP=<16|32|64|128>
#include "fgnn_final_tc.cuh"

#define FGNN_CAT2(a, b, c, d) a##b##c##d
#define FGNN_CAT(a, b, c, d) FGNN_CAT2(a, b, c, d)

namespace fgnn {
typedef void (*final_kernel_t)(Params);
typedef void (*dense_kernel_t)(const float*, const float*, float*, const float*, int, int);

final_kernel_t FGNN_CAT(get_final_k, FGNN_K, _hp, FGNN_HP)(bool closed) {
    return closed ? k_final<FGNN_K, FGNN_HP, true> : k_final<FGNN_K, FGNN_HP, false>;
}
dense_kernel_t FGNN_CAT(get_dense_k, FGNN_K, _hp, FGNN_HP)() { return k_actor_dense<FGNN_K, FGNN_HP>; }

typedef void (*final_tc_kernel_t)(Params, const uint8_t*);
final_tc_kernel_t FGNN_CAT(get_final_tc_k, FGNN_K, _hp, FGNN_HP)(bool closed) {
#if FGNN_HP <= 64
    return closed ? k_final_tc<FGNN_K, FGNN_HP, true> : k_final_tc<FGNN_K, FGNN_HP, false>;
#else
    (void)closed;
    return nullptr;      // HP = 128: operands do not fit shared memory; FFMA path only
#endif
}
}  // namespace fgnn
