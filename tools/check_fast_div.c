// Exhaustive host check (all 2^32 float32 bit patterns) of the division-free output scaling of audio_kernels.cu (scale_out):
//   q = u * y; r = fma(-q, c, u); q1 = fma(r, y, q) with c = -1090519040, y = RN(1/c)  ==  u / c  except for 0 < |u| < 2^-102 and inf
// (measured: 129,058 mismatches, the largest finite failing |u| = 0x1.fffffep-103).  gcc -O2 -fopenmp -ffp-contract=off check_fast_div.c -lm
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <omp.h>
int main(void) {
    const float c = -1090519040.0f;
    const float y = 1.0f / c;   // correctly rounded reciprocal
    uint64_t bad = 0, checked = 0;
    uint32_t first_bad = 0;
#pragma omp parallel for reduction(+:bad,checked) schedule(static)
    for (uint64_t i = 0; i < (1ull << 32); i++) {
        uint32_t b = (uint32_t)i;
        float x; memcpy(&x, &b, 4);
        if (isnan(x)) continue;
        volatile float want_v = x / c; float want = want_v;
        float q0 = x * y;
        float r = fmaf(-q0, c, x);
        float q1 = fmaf(r, y, q0);
        uint32_t a, w; memcpy(&a, &q1, 4); memcpy(&w, &want, 4);
        checked++;
        if (a != w) { bad++; if (!first_bad) first_bad = b; }
    }
    printf("y=%a checked=%llu bad=%llu first_bad=0x%08x\n", y, (unsigned long long)checked, (unsigned long long)bad, first_bad);
    return 0;
}
