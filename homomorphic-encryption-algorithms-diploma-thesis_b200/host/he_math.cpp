// he_math.cpp -- host mirror of src/core/he_math.cpp:22-269.  The reference spells every routine out as a
// chain of operator-DSL statements; here the recurring three-call idioms are named once (Chain below) and
// the routines read as their formulas.  The SEQUENCE of evaluator calls -- which constant is encoded at
// which ciphertext's parms_id and scale, when a level is burnt by a multiplication with an encoding of 1,
// where relinearize and rescale sit -- is the reference's, so the resulting ciphertexts are the ones the
// reference produces on SEAL.
#include "he_math.h"

#include <cmath>
#include <stdexcept>

#include "he_util.h"

namespace he::math {

namespace {

using he::gpu::Plaintext;

struct Chain {
    const CKKSEncoder &cencd;
    const Evaluator &eval;
    const RelinKeys &rk;

    // a constant encoded where `at` currently lives (its level and its exact scale)
    Plaintext constant(double value, const Ciphertext &at) const
    {
        Plaintext pt;
        cencd.encode(value, at.parms_id(), at.scale(), pt);
        return pt;
    }
    // ct <- rescale(ct * value): one level
    void scale_by(Ciphertext &ct, double value) const
    {
        eval.multiply_plain_inplace(ct, constant(value, ct));
        eval.rescale_to_next_inplace(ct);
    }
    void burn_level(Ciphertext &ct) const { scale_by(ct, 1.0); }
    void shift_by(Ciphertext &ct, double value) const { eval.add_plain_inplace(ct, constant(value, ct)); }
    // ct <- rescale(relin(ct * other)): one level
    void times(Ciphertext &ct, const Ciphertext &other) const
    {
        eval.multiply_inplace(ct, other);
        eval.relinearize_inplace(ct, rk);
        eval.rescale_to_next_inplace(ct);
    }
    void squared(Ciphertext &ct) const
    {
        eval.square_inplace(ct);
        eval.relinearize_inplace(ct, rk);
        eval.rescale_to_next_inplace(ct);
    }
};

}  // namespace

// 1/x = a * prod_{i>=0} (1 + (1 - a x)^(2^i)); the first two factors are folded into 2a - a^2 x
// (he_math.cpp:22-90)
Ciphertext signed_inv(const CKKSEncoder &cencd, const Evaluator &eval, const RelinKeys &rk, const Ciphertext &x_ct, double a,
                      std::size_t iter_num)
{
    if (iter_num == 0) throw std::invalid_argument("iter_num must be positive");
    const Chain ch{ cencd, eval, rk };
    Ciphertext y(x_ct);
    ch.scale_by(y, -a * a);  // -a^2 x
    ch.shift_by(y, 2 * a);   // 2a - a^2 x
    if (iter_num == 1) return y;

    Ciphertext t(x_ct);  // t = a x - 1, squared once per further factor
    ch.scale_by(t, a);
    const Plaintext one = ch.constant(1.0, t);
    eval.sub_plain_inplace(t, one);
    eval.multiply_plain_inplace(y, one);  // y joins t's next level
    eval.rescale_to_next_inplace(y);
    for (std::size_t it = 1; it < iter_num; ++it) {
        ch.squared(t);
        Ciphertext factor;
        eval.add_plain(t, ch.constant(1.0, t), factor);  // 1 + (a x - 1)^(2^it)
        ch.times(y, factor);
    }
    return y;
}

// Newton for y^-2 = 2x: y <- 3/2 y - x y^3, first step taken on the plaintext guess (he_math.cpp:95-205,
// the depth-2-per-iteration variant the reference compiles)
Ciphertext inv_sqrt_twice(const CKKSEncoder &cencd, const Evaluator &eval, const RelinKeys &rk, const Ciphertext &x_ct, double a,
                          std::size_t iter_num)
{
    if (iter_num == 0) throw std::invalid_argument("iter_num must be positive");
    const Chain ch{ cencd, eval, rk };
    Ciphertext x(x_ct), y(x_ct);
    ch.scale_by(y, -a * a * a);  // -a^3 x
    ch.shift_by(y, 1.5 * a);     // 3/2 a - a^3 x
    for (std::size_t it = 1; it < iter_num; ++it) {
        Ciphertext cube(y);  // becomes x y^3 two levels below y
        ch.scale_by(y, 1.5);
        ch.burn_level(y);
        for (int lvl = it > 1 ? 2 : 1; lvl > 0; --lvl) ch.burn_level(x);  // x down to the previous y
        Ciphertext xy(x);
        ch.times(xy, cube);  // x y
        ch.squared(cube);    // y^2
        ch.times(cube, xy);  // x y^3
        eval.sub_inplace(y, cube);
    }
    return y;
}

// sqrt(x) = (1/sqrt(2x)) * (sqrt(2) x)  (he_math.cpp:210-232)
Ciphertext sqrt(const SEALContext &ctx, const CKKSEncoder &cencd, const Evaluator &eval, const RelinKeys &rk, const Ciphertext &x_ct,
                double a, std::size_t iter_num)
{
    const Chain ch{ cencd, eval, rk };
    Ciphertext y = inv_sqrt_twice(cencd, eval, rk, x_ct, 1.0 / a / std::sqrt(2.0), iter_num);
    Ciphertext sx(x_ct);
    ch.scale_by(sx, std::sqrt(2.0));
    Plaintext scratch;
    he::util::reach_chain_level(ctx, cencd, eval, scratch, sx, y);
    ch.times(y, sx);
    return y;
}

// |x| = (1/sqrt(2 x^2)) * (sqrt(2) x^2)  (he_math.cpp:237-269)
Ciphertext abs(const SEALContext &ctx, const CKKSEncoder &cencd, const Evaluator &eval, const RelinKeys &rk, const Ciphertext &x_ct,
               double a, std::size_t iter_num)
{
    const Chain ch{ cencd, eval, rk };
    Ciphertext x2(x_ct);
    ch.squared(x2);
    Ciphertext y = inv_sqrt_twice(cencd, eval, rk, x2, 1.0 / a / std::sqrt(2.0), iter_num);
    ch.scale_by(x2, std::sqrt(2.0));
    for (std::size_t lvl = he::util::get_chain_index(ctx, x2) - he::util::get_chain_index(ctx, y); lvl > 0; --lvl) ch.burn_level(x2);
    ch.times(y, x2);
    return y;
}

}  // namespace he::math
