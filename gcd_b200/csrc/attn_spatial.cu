// placeholder until the tcgen05 flash-attention kernel lands (next commit)
#include "common.cuh"
#include "../../include/gcd_b200.h"
extern "C" int gcd_attention_spatial(const void*, int, int, int, void*, void*) {
    gcd_set_error("gcd_attention_spatial: not built yet");
    return -3;
}
