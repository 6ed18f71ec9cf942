#define __host__
#define __device__
#define __forceinline__ inline
#include "dropout.cuh"
#include <cstdio>
#include <cstdlib>
using namespace abcgpt;
int main(int argc, char** argv) {
  const uint32_t key = strtoul(argv[1], nullptr, 10);
  const int rows = atoi(argv[2]), cols = atoi(argv[3]);
  const float p = atof(argv[4]);
  const DropCfg d = make_drop(p, key);
  // residual / embedding sites
  for (int r = 0; r < rows; ++r) {
    const uint32_t rk = drop_row_key(d.key, r);
    for (int c = 0; c < cols; ++c) {
      const uint32_t bits = drop_pair_bits(rk, c >> 1);
      putchar(((c & 1) ? drop_keep_hi(bits, d.thr16) : drop_keep_lo(bits, d.thr16)) ? '1' : '0');
    }
  }
  putchar('\n');
  // attention-probability site
  for (int r = 0; r < rows; ++r) {
    const uint32_t a = drop_row_key(d.key, r), b = drop_row_key2(a);
    for (int c = 0; c < cols; ++c) {
      const uint32_t u = attn_drop_signs(attn_drop_fold(a + (uint32_t)(c >> 1) * kDropWeyl, b), d.k15);
      putchar(((c & 1) ? (u >> 31) : ((u >> 15) & 1)) ? '1' : '0');
    }
  }
  putchar('\n');
  return 0;
}
