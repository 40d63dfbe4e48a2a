/* test helper: array versions of include/vkp_detmath.h compiled for the host (gcc -ffp-contract=off) */
#include "../../include/vkp_detmath.h"
#define EXPORT __attribute__((visibility("default")))
/* func: 0 sin 1 cos 2 atan2(x,y) 3 asin 4 pow(x,y) 5 log2 6 exp2 7 f32->f16->f32 8 unorm8 */
EXPORT void probe_eval(int func, const float* x, const float* y, int n, float* out)
{
    for (int i = 0; i < n; i++) {
        switch (func) {
        case 0: out[i] = vkp_sinf(x[i]); break;
        case 1: out[i] = vkp_cosf(x[i]); break;
        case 2: out[i] = vkp_atan2f(x[i], y[i]); break;
        case 3: out[i] = vkp_asinf(x[i]); break;
        case 4: out[i] = vkp_powf(x[i], y[i]); break;
        case 5: out[i] = vkp_log2f(x[i]); break;
        case 6: out[i] = vkp_exp2f(x[i]); break;
        case 7: out[i] = vkp_f16_to_f32(vkp_f32_to_f16(x[i])); break;
        default: out[i] = (float)vkp_unorm8(x[i]); break;
        }
    }
}
