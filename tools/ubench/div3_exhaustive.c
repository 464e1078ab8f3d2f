#include <stdio.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
int main() {
    const float y = 1.0f / 3.0f;
    uint64_t bad = 0, n = 0;
    // all positive floats from the smallest normal up to 8.0 (and a band of subnormals)
    for (uint32_t u = 0x00000001u; u <= 0x7F000000u; ++u) {
        float a; memcpy(&a, &u, 4);
        float q = a * y;
        float r = fmaf(-3.0f, q, a);
        float q2 = fmaf(r, y, q);
        float ref = a / 3.0f;
        ++n;
        if (q2 != ref) { if (bad < 5) printf("mismatch a=%a q2=%a ref=%a\n", a, q2, ref); ++bad; }
    }
    printf("tested %llu values, %llu mismatches\n", (unsigned long long)n, (unsigned long long)bad);
    return 0;
}
