// Warp-level tensor-core product on TF32 hi/lo splits ("3xTF32": x*y ~ xh*yh + xh*yl + xl*yh, fp32 accumulate,
// ~fp32 accuracy), used where a small reduction over bins rides along with other work of the warp: the anchor
// estimator's weighted sums (attractor.cu) and the same sums inside the output projection's epilogue (gemm_tc.cu).
#pragma once
#include <stdint.h>

namespace danet {

// round to the 10-bit TF32 mantissa: add half an ulp of the kept part, clear the dropped 13 bits (round half away from
// zero in magnitude, as cvt.rna.tf32; two integer instructions where the cvt expands to a NaN-safe sequence of eight --
// the operands here are finite by construction)
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// x = hi + lo with hi exactly representable in TF32; the tensor core ignores the low 13 bits of lo (a 2^-21 effect)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// D[16x8] += A[16x8] * B[8x8]; fragments as in the PTX ISA (gid = lane / 4, tig = lane % 4):
//   a0 (row gid, k tig), a1 (row gid + 8, k tig), a2 (row gid, k tig + 4), a3 (row gid + 8, k tig + 4)
//   b0 (k tig, col gid), b1 (k tig + 4, col gid)
//   d0 (row gid, col 2 tig), d1 (row gid, col 2 tig + 1), d2 / d3 the same for row gid + 8
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

}  // namespace danet
