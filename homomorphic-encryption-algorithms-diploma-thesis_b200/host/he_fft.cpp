// he_fft.cpp -- homomorphic FFT on the GPU evaluator (mirror of src/core/he_fft.cpp).
//
// fft_: the reference recursion (he_fft.cpp:13-68) is restated bottom-up: at every recursion
// level all n/2 butterflies  t = rescale(odd*w^k), e = rescale(even*1), (e+t, e-t)  are
// independent, so one level is ONE batched device call (hegpu_fft_butterflies) instead of
// n/2 x 6 evaluator calls.  Plaintexts (w^k, 1) are encoded on the host exactly where the
// reference encodes them (he_fft.cpp:47,55).
// bfft_: log2(n) stages y <- D0*y + D1*rot(y,+s) [+ D2*rot(y,-s)], each rescaled
// (he_fft.cpp:178-203) = one hegpu_bfft_stage per stage; the three stage diagonals are built
// and encoded on the host as in diag_D (he_fft.cpp:89-164).
#include "he_fft.h"

#include <cmath>

#include "he_operators.h"

namespace he::fft {

using he::gpu::Plaintext;
using Complex = std::complex<double>;

std::vector<Plaintext> *encoded_log = nullptr;

namespace {

void log_pt(const Plaintext &p)
{
    if (encoded_log) encoded_log->push_back(p);
}

// gather single ciphertexts into a device batch / scatter a batch back
struct Batch {
    hegpu_ct *h = nullptr;
    ~Batch() { hegpu_ct_destroy(h); }
};
void gather(const he::gpu::SEALContext &ctx, Batch &b, const std::vector<const Ciphertext *> &src, std::uint32_t size_cap)
{
    he::gpu::check(hegpu_ct_create(ctx.raw(), &b.h, (std::uint32_t)src.size(), size_cap, ctx.key_limbs() - 1));
    for (std::size_t i = 0; i < src.size(); ++i) he::gpu::check(hegpu_ct_copy_one(ctx.raw(), b.h, (std::uint32_t)i, src[i]->handle(), 0));
}
void scatter(const he::gpu::SEALContext &ctx, const Batch &b, std::vector<Ciphertext> &dst)
{
    for (std::size_t i = 0; i < dst.size(); ++i) {
        hegpu_ct *d = dst[i].prepare(ctx);
        he::gpu::check(hegpu_ct_copy_one(ctx.raw(), d, 0, b.h, (std::uint32_t)i));
    }
}

// One recursion level combining `even` and `odd` sub-transforms of length n/2 each
// (he_fft.cpp:33-66); sign = +1 forward, -1 inverse twiddles.
std::vector<Ciphertext> combine(const CKKSEncoder &cencd, const Evaluator &eval, const std::vector<Ciphertext> &even,
                                const std::vector<Ciphertext> &odd, int sign)
{
    const auto &ctx = eval.context();
    const std::size_t half = even.size(), n = 2 * half;
    const double scale = even[0].scale();
    const auto pid = even[0].parms_id();
    const Complex w = std::exp(Complex(0, sign * -2 * M_PI / (double)n));

    Plaintext one_pt;
    cencd.encode(Complex(1, 0), pid, scale, one_pt);
    log_pt(one_pt);
    // w^k plaintext set, k < n/2 (encoded one by one on the host, uploaded as one set)
    const std::size_t per = (std::size_t)pid.limbs * ctx.poly_modulus_degree();
    std::vector<std::uint64_t> all(half * per);
    for (std::size_t k = 0; k < half; ++k) {
        Plaintext wk;
        cencd.encode(std::pow(w, (double)k), pid, scale, wk);
        log_pt(wk);
        const std::vector<std::uint64_t> limbs = wk.save();
        std::copy(limbs.begin(), limbs.end(), all.begin() + k * per);
    }
    hegpu_pt *wset = nullptr;
    he::gpu::check(hegpu_pt_create(ctx.raw(), &wset, (std::uint32_t)half, pid.limbs));
    std::shared_ptr<hegpu_pt> wguard(wset, [](hegpu_pt *p) { hegpu_pt_destroy(p); });
    he::gpu::check(hegpu_pt_upload(wset, all.data(), pid.limbs, scale));

    std::vector<const Ciphertext *> ev, od;
    for (auto &c : even) ev.push_back(&c);
    for (auto &c : odd) od.push_back(&c);
    Batch be, bo, bout;
    gather(ctx, be, ev, 2);
    gather(ctx, bo, od, 2);
    he::gpu::check(hegpu_ct_create(ctx.raw(), &bout.h, (std::uint32_t)n, 2, ctx.key_limbs() - 1));
    he::gpu::check(hegpu_fft_butterflies(ctx.raw(), bout.h, be.h, bo.h, wset, one_pt.handle()));
    std::vector<Ciphertext> res(n);
    scatter(ctx, bout, res);
    return res;
}

std::vector<Ciphertext> fft_rec(const CKKSEncoder &cencd, const Evaluator &eval, const std::vector<Ciphertext> &v, int sign)
{
    const std::size_t n = v.size();
    if (n == 1) return v;
    std::vector<Ciphertext> e(n / 2), o(n / 2);
    for (std::size_t i = 0; i < n / 2; ++i) {
        e[i] = v[2 * i];
        o[i] = v[2 * i + 1];
    }
    return combine(cencd, eval, fft_rec(cencd, eval, e, sign), fft_rec(cencd, eval, o, sign), sign);
}

// the three diagonals of stage matrix k = 2^i (he_fft.cpp:89-164), tiled over all slots
std::vector<Complex> stage_diagonal(int d, std::size_t k, std::size_t n, std::size_t slots, int sign)
{
    const std::size_t blk = n / k;  // run length of ones / twiddles
    const Complex w = std::exp(Complex(0, sign * -2 * M_PI / (double)(2 * blk)));
    std::vector<Complex> diag(n, Complex(0, 0));
    for (std::size_t pos = 0; pos < n; ++pos) {
        const std::size_t run = pos / blk, off = pos % blk;  // alternating runs of length blk
        const bool odd_run = run & 1;
        const bool last_run = run == k - 1, first_run = run == 0;
        if (d == 0) diag[pos] = odd_run ? -std::pow(w, (double)off) : Complex(1, 0);
        if (d == 1) diag[pos] = !odd_run ? Complex(1, 0) : ((k == 2 && last_run) ? std::pow(w, (double)off) : Complex(0, 0));
        if (d == 2) diag[pos] = (odd_run && !first_run) ? std::pow(w, (double)off) : Complex(0, 0);
    }
    std::vector<Complex> rep;
    rep.reserve(slots);
    for (std::size_t r = 0; r < slots / n; ++r) rep.insert(rep.end(), diag.begin(), diag.end());
    return rep;
}

Ciphertext bfft_impl(const CKKSEncoder &cencd, const Evaluator &eval, const Ciphertext &x_ct, std::size_t n, int sign)
{
    const auto &ctx = eval.context();
    Ciphertext y(x_ct);
    std::size_t stages = 0;
    while ((std::size_t(1) << stages) < n) ++stages;
    const std::size_t per = ctx.poly_modulus_degree();
    for (std::size_t i = 1; i <= stages; ++i) {
        const std::size_t k = std::size_t(1) << i;
        const int steps = (int)(n / k);
        const bool with_d2 = i != 1;
        const auto pid = y.parms_id();
        const double scale = y.scale();
        const std::uint32_t cnt = with_d2 ? 3 : 2;
        std::vector<std::uint64_t> all((std::size_t)cnt * pid.limbs * per);
        for (std::uint32_t d = 0; d < cnt; ++d) {
            Plaintext pt;
            cencd.encode(stage_diagonal((int)d, k, n, cencd.slot_count(), sign), pid, scale, pt);
            log_pt(pt);
            const auto limbs = pt.save();
            std::copy(limbs.begin(), limbs.end(), all.begin() + (std::size_t)d * pid.limbs * per);
        }
        hegpu_pt *set = nullptr;
        he::gpu::check(hegpu_pt_create(ctx.raw(), &set, cnt, pid.limbs));
        std::shared_ptr<hegpu_pt> guard(set, [](hegpu_pt *p) { hegpu_pt_destroy(p); });
        he::gpu::check(hegpu_pt_upload(set, all.data(), pid.limbs, scale));
        he::gpu::check(hegpu_bfft_stage(ctx.raw(), y.handle(), set, steps, with_d2 ? 1 : 0));
    }
    return y;
}

}  // namespace

std::vector<Ciphertext> fft(const CKKSEncoder &cencd, const Evaluator &eval, const std::vector<Ciphertext> &vec_ct)
{
    return fft_rec(cencd, eval, vec_ct, +1);
}

std::vector<Ciphertext> ifft(const CKKSEncoder &cencd, const Evaluator &eval, const std::vector<Ciphertext> &vec_ct)
{
    using he::operators::operator%;
    std::vector<Ciphertext> res = fft_rec(cencd, eval, vec_ct, -1);
    Plaintext n_inv_pt;
    cencd.encode(1.0 / (double)vec_ct.size(), res[0].parms_id(), res[0].scale(), n_inv_pt);
    log_pt(n_inv_pt);
    for (auto &r : res) {
        he::operators::operator*=(r, eval % n_inv_pt);
        he::operators::operator^=(r, eval);
    }
    return res;
}

Ciphertext bfft(const CKKSEncoder &cencd, const Evaluator &eval, const GaloisKeys &, const Ciphertext &x_ct, std::size_t n)
{
    return bfft_impl(cencd, eval, x_ct, n, +1);
}

Ciphertext ibfft(const CKKSEncoder &cencd, const Evaluator &eval, const GaloisKeys &, const Ciphertext &x_ct, std::size_t n)
{
    using he::operators::operator%;
    Ciphertext res = bfft_impl(cencd, eval, x_ct, n, -1);
    Plaintext n_inv_pt;
    cencd.encode(1.0 / (double)n, res.parms_id(), res.scale(), n_inv_pt);
    log_pt(n_inv_pt);
    he::operators::operator*=(res, eval % n_inv_pt);
    he::operators::operator^=(res, eval);
    return res;
}

}  // namespace he::fft
