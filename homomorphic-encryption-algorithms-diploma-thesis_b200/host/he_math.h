// he_math.h -- host mirror of the reference's include/he_math.h:1-37 ("next" N3): Newton /
// product iterations for 1/x, 1/sqrt(2x), sqrt(x) and |x| on encrypted vectors.  Same names, argument
// order and meaning; every step is one evaluator call of the C ABI, so the ciphertext never leaves
// the device.  `a` is the caller's initial guess (conditions as in the reference header).
#pragma once
#include <cstddef>

#include "hegpu_seal_like.hpp"

namespace he::math {

using he::gpu::Ciphertext;
using he::gpu::CKKSEncoder;
using he::gpu::Evaluator;
using he::gpu::RelinKeys;
using he::gpu::SEALContext;

// f(x) = 1/x; needs |a*x - 1| < 1.  Consumes iter_num + 1 levels (1 level when iter_num == 1).
Ciphertext signed_inv(const CKKSEncoder &cencd, const Evaluator &eval, const RelinKeys &rk, const Ciphertext &x_ct, double a,
                      std::size_t iter_num);
// f(x) = 1/sqrt(2x), x > 0; needs 0 < a < sqrt(3/(2x)).  Consumes 1 + 2*(iter_num - 1) levels.
Ciphertext inv_sqrt_twice(const CKKSEncoder &cencd, const Evaluator &eval, const RelinKeys &rk, const Ciphertext &x_ct, double a,
                          std::size_t iter_num);
// f(x) = sqrt(x) = (1/sqrt(2x)) * (sqrt(2) x); `a` is a guess of sqrt(x)
Ciphertext sqrt(const SEALContext &ctx, const CKKSEncoder &cencd, const Evaluator &eval, const RelinKeys &rk, const Ciphertext &x_ct,
                double a, std::size_t iter_num);
// f(x) = |x| = (1/sqrt(2x^2)) * (sqrt(2) x^2); `a` is a guess of |x|
Ciphertext abs(const SEALContext &ctx, const CKKSEncoder &cencd, const Evaluator &eval, const RelinKeys &rk, const Ciphertext &x_ct,
               double a, std::size_t iter_num);

}  // namespace he::math
