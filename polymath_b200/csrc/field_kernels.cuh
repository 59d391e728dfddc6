// Elementwise Fr/Fq batch kernels and the INT32 IMAD-pipe peak microbenchmark (interface).
#pragma once
#include "common.cuh"
#include "field.cuh"

namespace pm {

enum class FieldOp { Mul, Add, Sub, Inv };   // Inv: out = a^-1 (b ignored)
void launch_fr_batch(FieldOp op, const Fr* a, const Fr* b, Fr* out, size_t n, cudaStream_t stream);
void launch_fq_batch(FieldOp op, const Fq* a, const Fq* b, Fq* out, size_t n, cudaStream_t stream);

// Dependency-free IMAD.WIDE.U32 issue-rate microbenchmark: returns 32x32+64 multiply-adds per second.
double measure_imad_peak(int iters);
// Register-resident Montgomery products per second (chains of `depth` dependent products per thread).
double measure_fr_mul_rate(int depth);
double measure_fq_mul_rate(int depth);
double measure_fq_variant_rate(int mode, int depth);   // 1 = squaring, 2 = Karatsuba product
double measure_fr_variant_rate(int mode, int depth);

}  // namespace pm
