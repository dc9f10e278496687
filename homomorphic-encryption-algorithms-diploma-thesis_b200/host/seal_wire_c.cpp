// seal_wire_c.cpp -- plain C entry points of the wire layer (libseal_wire.so, CPU only) so that tests and other
// languages can drive it without the C++ types.  Every function returns 0 on success; hewire_last_error() holds the
// message of the last failure on this thread.  Buffers returned through `**` are malloc'ed: release with hewire_free.
#include <cstdlib>
#include <cstring>
#include <string>

#include "seal_wire.hpp"

using namespace he::wire;

namespace {
thread_local std::string g_err;
template <class F>
int guard(F &&f)
{
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
std::uint8_t *dup(const void *p, std::size_t n)
{
    std::uint8_t *m = static_cast<std::uint8_t *>(std::malloc(n ? n : 1));
    if (n) std::memcpy(m, p, n);
    return m;
}
}  // namespace

extern "C" {

const char *hewire_last_error() { return g_err.c_str(); }
void hewire_free(void *p) { std::free(p); }
int hewire_zstd_available() { return zstd_available() ? 1 : 0; }

int hewire_blake2b(std::uint8_t *out, std::size_t outlen, const void *in, std::size_t inlen, const void *key, std::size_t keylen)
{
    return guard([&] { blake2b(out, outlen, in, inlen, key, keylen); });
}
int hewire_blake2xb(std::uint8_t *out, std::size_t outlen, const void *in, std::size_t inlen, const void *key, std::size_t keylen)
{
    return guard([&] { blake2xb(out, outlen, in, inlen, key, keylen); });
}
int hewire_prng_generate(const std::uint64_t seed[8], std::size_t count, std::uint8_t *dst)
{
    return guard([&] {
        std::array<std::uint64_t, 8> s;
        std::memcpy(s.data(), seed, 64);
        Blake2xbPrng prng(s);
        prng.generate(count, dst);
    });
}
int hewire_sample_poly_uniform(const std::uint64_t seed[8], const std::uint64_t *moduli, std::size_t limbs, std::size_t n, std::uint64_t *dst)
{
    return guard([&] {
        std::array<std::uint64_t, 8> s;
        std::memcpy(s.data(), seed, 64);
        Blake2xbPrng prng(s);
        sample_poly_uniform(prng, std::vector<std::uint64_t>(moduli, moduli + limbs), n, dst);
    });
}
int hewire_unwrap(const std::uint8_t *in, std::size_t avail, std::size_t *consumed, std::uint8_t **members, std::size_t *members_len)
{
    return guard([&] {
        const bytes m = unwrap(in, avail, consumed);
        *members = dup(m.data(), m.size());
        *members_len = m.size();
    });
}
int hewire_wrap(const std::uint8_t *members, std::size_t len, int mode, std::uint8_t **out, std::size_t *out_len)
{
    return guard([&] {
        const bytes w = wrap(bytes(members, members + len), (compr_mode)mode);
        *out = dup(w.data(), w.size());
        *out_len = w.size();
    });
}
// parameters: moduli must hold 64 words
int hewire_load_parms(const std::uint8_t *in, std::size_t avail, std::size_t *consumed, std::uint8_t *scheme, std::uint64_t *n, std::uint64_t *k,
                      std::uint64_t *moduli, std::uint64_t *plain_modulus)
{
    return guard([&] {
        Parms p;
        *consumed = load_parms(in, avail, p);
        *scheme = p.scheme;
        *n = p.n;
        *k = p.moduli.size();
        std::memcpy(moduli, p.moduli.data(), p.moduli.size() * 8);
        *plain_modulus = p.plain_modulus;
    });
}
int hewire_save_parms(std::uint64_t n, const std::uint64_t *moduli, std::size_t k, int mode, std::uint8_t **out, std::size_t *out_len)
{
    return guard([&] {
        Parms p;
        p.n = n;
        p.moduli.assign(moduli, moduli + k);
        const bytes w = save_parms(p, (compr_mode)mode);
        *out = dup(w.data(), w.size());
        *out_len = w.size();
    });
}
int hewire_parms_id(std::uint64_t n, const std::uint64_t *moduli, std::size_t k, std::size_t limbs, std::uint64_t id[4])
{
    return guard([&] {
        Parms p;
        p.n = n;
        p.moduli.assign(moduli, moduli + k);
        const parms_id_t i = p.parms_id(limbs);
        std::memcpy(id, i.data(), 32);
    });
}
// ciphertext at `in` under the chain (n, moduli[k]); *data = [size][limbs][n]
int hewire_load_ciphertext(std::uint64_t n, const std::uint64_t *moduli, std::size_t k, const std::uint8_t *in, std::size_t avail, std::size_t *consumed,
                           std::uint64_t *size, std::uint64_t *limbs, double *scale, int *seeded, std::uint64_t **data)
{
    return guard([&] {
        Parms p;
        p.n = n;
        p.moduli.assign(moduli, moduli + k);
        CtData c;
        *consumed = load_ciphertext(p, in, avail, c);
        *size = c.size;
        *limbs = c.limbs;
        *scale = c.scale;
        *seeded = c.was_seeded ? 1 : 0;
        *data = reinterpret_cast<std::uint64_t *>(dup(c.data.data(), c.data.size() * 8));
    });
}
// seed == NULL: plain form; otherwise the seeded form (the caller guarantees data(1) = expansion of the seed)
int hewire_save_ciphertext(std::uint64_t n, const std::uint64_t *moduli, std::size_t k, std::uint64_t size, std::uint64_t limbs, double scale,
                           const std::uint64_t *data, const std::uint64_t *seed, int mode, std::uint8_t **out, std::size_t *out_len)
{
    return guard([&] {
        Parms p;
        p.n = n;
        p.moduli.assign(moduli, moduli + k);
        CtData c;
        c.size = size;
        c.limbs = limbs;
        c.n = n;
        c.scale = scale;
        c.parms_id = p.parms_id((std::size_t)limbs);
        c.data.assign(data, data + size * limbs * n);
        std::array<std::uint64_t, 8> s;
        if (seed) std::memcpy(s.data(), seed, 64);
        const bytes w = save_ciphertext(c, (compr_mode)mode, seed ? &s : nullptr);
        *out = dup(w.data(), w.size());
        *out_len = w.size();
    });
}
// key-switching keys: *flat = index 0 flattened to [digits][2][k][n]
int hewire_load_kswitch_keys(std::uint64_t n, const std::uint64_t *moduli, std::size_t k, const std::uint8_t *in, std::size_t avail, std::size_t *consumed,
                             std::uint64_t *indices, std::uint64_t *digits, std::uint64_t **flat)
{
    return guard([&] {
        Parms p;
        p.n = n;
        p.moduli.assign(moduli, moduli + k);
        KSwitchData d;
        *consumed = load_kswitch_keys(p, in, avail, d);
        *indices = d.keys.size();
        *digits = d.keys.empty() ? 0 : d.keys[0].size();
        const auto f = d.keys.empty() ? std::vector<std::uint64_t>() : d.flat(0);
        *flat = reinterpret_cast<std::uint64_t *>(dup(f.data(), f.size() * 8));
    });
}
// one index (RelinKeys): key [digits][2][k][n]; seeds [digits][8] or NULL
int hewire_save_kswitch_keys(std::uint64_t n, const std::uint64_t *moduli, std::size_t k, const std::uint64_t *key, std::size_t digits,
                             const std::uint64_t *seeds, int mode, std::uint8_t **out, std::size_t *out_len)
{
    return guard([&] {
        Parms p;
        p.n = n;
        p.moduli.assign(moduli, moduli + k);
        KSwitchData d;
        d.parms_id = p.key_parms_id();
        d.keys.resize(1);
        std::vector<std::vector<std::array<std::uint64_t, 8>>> sd(1);
        const std::size_t words = 2 * k * n;
        for (std::size_t j = 0; j < digits; ++j) {
            CtData c;
            c.size = 2;
            c.limbs = k;
            c.n = n;
            c.scale = 1.0;
            c.parms_id = d.parms_id;
            c.data.assign(key + j * words, key + (j + 1) * words);
            d.keys[0].push_back(c);
            if (seeds) {
                std::array<std::uint64_t, 8> s;
                std::memcpy(s.data(), seeds + j * 8, 64);
                sd[0].push_back(s);
            }
        }
        const bytes w = save_kswitch_keys(d, (compr_mode)mode, seeds ? &sd : nullptr);
        *out = dup(w.data(), w.size());
        *out_len = w.size();
    });
}
}
