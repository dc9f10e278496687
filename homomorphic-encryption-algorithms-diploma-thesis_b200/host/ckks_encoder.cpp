// CKKSEncoder stand-in: conjugate-extend the slot vector, inverse negacyclic
// embedding in double precision, scale, round, RNS-decompose, NTT (on the GPU).
#include <cmath>

#include "hegpu_seal_like.hpp"

namespace he::gpu {

namespace {
void fft_inplace(std::vector<std::complex<double>> &a)  // forward DFT, radix-2, size = power of two
{
    const std::size_t n = a.size();
    for (std::size_t i = 1, j = 0; i < n; ++i) {
        std::size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (std::size_t len = 2; len <= n; len <<= 1) {
        const double ang = -2.0 * M_PI / (double)len;
        const std::complex<double> wl(std::cos(ang), std::sin(ang));
        for (std::size_t i = 0; i < n; i += len) {
            std::complex<double> w(1.0, 0.0);
            for (std::size_t k = 0; k < len / 2; ++k) {
                // recompute the twiddle from its angle every 64 steps to bound drift
                if ((k & 63) == 0) w = std::complex<double>(std::cos(ang * (double)k), std::sin(ang * (double)k));
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
                w *= wl;
            }
        }
    }
}
}  // namespace

CKKSEncoder::CKKSEncoder(const SEALContext &ctx) : ctx_(ctx), n_(ctx.poly_modulus_degree()), slots_(n_ / 2)
{
    const std::uint64_t m = 2 * (std::uint64_t)n_;
    std::uint64_t pos = 1;
    idx1_.resize(slots_);
    idx2_.resize(slots_);
    for (std::size_t i = 0; i < slots_; ++i) {
        idx1_[i] = (std::uint32_t)((pos - 1) >> 1);
        idx2_[i] = (std::uint32_t)((m - pos - 1) >> 1);
        pos = pos * 3 % m;
    }
    zeta_neg_.resize(n_);
    for (std::size_t k = 0; k < n_; ++k) zeta_neg_[k] = std::polar(1.0, -M_PI * (double)k / (double)n_);
}

void CKKSEncoder::encode(const std::vector<std::complex<double>> &values, parms_id_type parms_id, double scale, Plaintext &dst) const
{
    if (values.size() > slots_) throw std::invalid_argument("values has invalid size");
    if (parms_id.limbs == 0 || parms_id.limbs > ctx_.key_limbs()) throw std::invalid_argument("parms_id is not valid for encryption parameters");
    std::vector<std::complex<double>> v(n_, { 0.0, 0.0 });
    for (std::size_t i = 0; i < values.size(); ++i) {
        v[idx1_[i]] = values[i];
        v[idx2_[i]] = std::conj(values[i]);
    }
    fft_inplace(v);  // m_k = zeta^-k * DFT(v)[k] / N
    std::vector<std::uint64_t> limbs((std::size_t)parms_id.limbs * n_);
    const auto &q = ctx_.coeff_modulus();
    for (std::size_t k = 0; k < n_; ++k) {
        const double c = std::nearbyint((v[k] * zeta_neg_[k]).real() / (double)n_ * scale);
        if (std::fabs(c) >= 0x1p62) throw std::invalid_argument("encoded values are too large");
        const long long ci = (long long)c;
        for (std::uint32_t i = 0; i < parms_id.limbs; ++i) {
            const long long qi = (long long)q[i];
            long long r = ci % qi;
            if (r < 0) r += qi;
            limbs[(std::size_t)i * n_ + k] = (std::uint64_t)r;
        }
    }
    check(hegpu_ntt_forward_host(ctx_.raw(), limbs.data(), parms_id.limbs, 0, parms_id.limbs));
    dst.load(ctx_, limbs.data(), parms_id.limbs, scale);
}

void CKKSEncoder::encode(const std::vector<double> &values, parms_id_type parms_id, double scale, Plaintext &dst) const
{
    std::vector<std::complex<double>> c(values.begin(), values.end());
    encode(c, parms_id, scale, dst);
}

void CKKSEncoder::encode(double value, parms_id_type parms_id, double scale, Plaintext &dst) const
{
    if (parms_id.limbs == 0 || parms_id.limbs > ctx_.key_limbs()) throw std::invalid_argument("parms_id is not valid for encryption parameters");
    const double c = std::nearbyint(value * scale);
    if (std::fabs(c) >= 0x1p62) throw std::invalid_argument("encoded value is too large");
    const long long ci = (long long)c;
    std::vector<std::uint64_t> limbs((std::size_t)parms_id.limbs * n_);
    const auto &q = ctx_.coeff_modulus();
    for (std::uint32_t i = 0; i < parms_id.limbs; ++i) {
        long long r = ci % (long long)q[i];
        if (r < 0) r += (long long)q[i];
        for (std::size_t k = 0; k < n_; ++k) limbs[(std::size_t)i * n_ + k] = (std::uint64_t)r;
    }
    dst.load(ctx_, limbs.data(), parms_id.limbs, scale);
}

}  // namespace he::gpu
