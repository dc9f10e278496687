// seal_wire.cpp -- see seal_wire.hpp.  Restates SEAL 4.1's serialization (format fidelity unpinned against real SEAL).
#include "seal_wire.hpp"

#include <algorithm>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <zlib.h>

namespace he::wire {

namespace {

// ------------------------------------------------------------------ BLAKE2b (RFC 7693)
const std::uint64_t IV[8] = { 0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                              0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL };
const std::uint8_t SIGMA[12][16] = {
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 },
    { 11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4 }, { 7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8 },
    { 9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13 }, { 2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9 },
    { 12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11 }, { 13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10 },
    { 6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5 }, { 10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0 },
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 }
};
inline std::uint64_t rotr(std::uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }
inline std::uint64_t load64(const std::uint8_t *p)
{
    std::uint64_t v;
    std::memcpy(&v, p, 8);  // little-endian host
    return v;
}

struct B2State {
    std::uint64_t h[8], t[2] = { 0, 0 };
    std::uint8_t buf[128];
    std::size_t buflen = 0, outlen = 0;

    void compress(const std::uint8_t *block, bool last)
    {
        std::uint64_t m[16], v[16];
        for (int i = 0; i < 16; ++i) m[i] = load64(block + 8 * i);
        for (int i = 0; i < 8; ++i) {
            v[i] = h[i];
            v[i + 8] = IV[i];
        }
        v[12] ^= t[0];
        v[13] ^= t[1];
        if (last) v[14] = ~v[14];
        auto G = [&](int r, int i, int a, int b, int c, int d) {
            v[a] = v[a] + v[b] + m[SIGMA[r][2 * i]];
            v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + m[SIGMA[r][2 * i + 1]];
            v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; ++r) {
            G(r, 0, 0, 4, 8, 12);
            G(r, 1, 1, 5, 9, 13);
            G(r, 2, 2, 6, 10, 14);
            G(r, 3, 3, 7, 11, 15);
            G(r, 4, 0, 5, 10, 15);
            G(r, 5, 1, 6, 11, 12);
            G(r, 6, 2, 7, 8, 13);
            G(r, 7, 3, 4, 9, 14);
        }
        for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
    }
    void init(const std::uint8_t param[64])
    {
        for (int i = 0; i < 8; ++i) h[i] = IV[i] ^ load64(param + 8 * i);
        outlen = param[0];
    }
    void update(const void *in_, std::size_t n)
    {
        const std::uint8_t *in = static_cast<const std::uint8_t *>(in_);
        while (n) {
            if (buflen == 128) {  // a full buffer is only compressed once more input shows that it is not the last block
                t[0] += 128;
                if (t[0] < 128) ++t[1];
                compress(buf, false);
                buflen = 0;
            }
            const std::size_t take = std::min<std::size_t>(n, 128 - buflen);
            std::memcpy(buf + buflen, in, take);
            buflen += take;
            in += take;
            n -= take;
        }
    }
    void final(std::uint8_t *out)
    {
        t[0] += buflen;
        if (t[0] < buflen) ++t[1];
        std::memset(buf + buflen, 0, 128 - buflen);
        compress(buf, true);
        std::uint8_t full[64];
        std::memcpy(full, h, 64);
        std::memcpy(out, full, outlen);
    }
};

void put32(std::uint8_t *p, std::uint32_t v) { std::memcpy(p, &v, 4); }

}  // namespace

void blake2b_param(std::uint8_t *out, std::size_t outlen, const std::uint8_t param[64], const void *in, std::size_t inlen, const void *key,
                   std::size_t keylen)
{
    if (outlen == 0 || outlen > 64 || keylen > 64 || param[0] != outlen || param[1] != keylen) throw std::invalid_argument("blake2b parameters");
    B2State s;
    s.init(param);
    if (keylen) {
        std::uint8_t block[128] = { 0 };
        std::memcpy(block, key, keylen);
        s.update(block, 128);
    }
    s.update(in, inlen);
    s.final(out);
}

void blake2b(std::uint8_t *out, std::size_t outlen, const void *in, std::size_t inlen, const void *key, std::size_t keylen)
{
    std::uint8_t P[64] = { 0 };
    P[0] = (std::uint8_t)outlen;
    P[1] = (std::uint8_t)keylen;
    P[2] = 1;  // fanout
    P[3] = 1;  // depth
    blake2b_param(out, outlen, P, in, inlen, key, keylen);
}

// BLAKE2X (blake2xb-ref.c): root hash H0 with the XOF length in the parameter block, then output block i =
// BLAKE2b(H0) with fanout = depth = 0, leaf_length = inner_length = 64, node_offset = i, digest_length = block size
void blake2xb(std::uint8_t *out, std::size_t outlen, const void *in, std::size_t inlen, const void *key, std::size_t keylen)
{
    if (outlen == 0 || outlen > 0xFFFFFFFFull) throw std::invalid_argument("blake2xb output length");
    std::uint8_t P[64] = { 0 };
    P[0] = 64;
    P[1] = (std::uint8_t)keylen;
    P[2] = 1;
    P[3] = 1;
    put32(P + 12, (std::uint32_t)outlen);  // xof_length
    std::uint8_t root[64];
    blake2b_param(root, 64, P, in, inlen, key, keylen);
    P[1] = 0;
    P[2] = 0;
    P[3] = 0;
    put32(P + 4, 64);  // leaf_length
    P[16] = 0;         // node_depth
    P[17] = 64;        // inner_length
    for (std::uint32_t i = 0; outlen > 0; ++i) {
        const std::size_t block = std::min<std::size_t>(outlen, 64);
        P[0] = (std::uint8_t)block;
        put32(P + 8, i);  // node_offset
        blake2b_param(out + (std::size_t)i * 64, block, P, root, 64, nullptr, 0);
        outlen -= block;
    }
}

void Blake2xbPrng::generate(std::size_t count, std::uint8_t *dst)
{
    while (count) {
        if (pos_ == buf_.size()) {
            blake2xb(buf_.data(), buf_.size(), &counter_, sizeof(counter_), seed_.data(), seed_.size() * sizeof(std::uint64_t));
            ++counter_;
            pos_ = 0;
        }
        const std::size_t take = std::min(count, buf_.size() - pos_);
        std::memcpy(dst, buf_.data() + pos_, take);
        pos_ += take;
        dst += take;
        count -= take;
    }
}

void sample_poly_uniform(Blake2xbPrng &prng, const std::vector<std::uint64_t> &moduli, std::size_t n, std::uint64_t *dst)
{
    prng.generate(moduli.size() * n * sizeof(std::uint64_t), reinterpret_cast<std::uint8_t *>(dst));
    constexpr std::uint64_t max_random = 0xFFFFFFFFFFFFFFFFull;
    for (std::size_t j = 0; j < moduli.size(); ++j) {
        const std::uint64_t q = moduli[j], max_multiple = max_random - max_random % q - 1;
        for (std::size_t i = 0; i < n; ++i) {
            std::uint64_t r = dst[j * n + i];
            while (r >= max_multiple) prng.generate(sizeof(r), reinterpret_cast<std::uint8_t *>(&r));
            dst[j * n + i] = r % q;
        }
    }
}

// ------------------------------------------------------------------ compression back ends
namespace {
struct ZBuf {
    const void *src;
    std::size_t size, pos;
};
struct ZOut {
    void *dst;
    std::size_t size, pos;
};
struct Zstd {
    void *lib = nullptr;
    void *(*createDStream)() = nullptr;
    std::size_t (*freeDStream)(void *) = nullptr;
    std::size_t (*decompressStream)(void *, ZOut *, ZBuf *) = nullptr;
    unsigned (*isError)(std::size_t) = nullptr;
    std::size_t (*compress)(void *, std::size_t, const void *, std::size_t, int) = nullptr;
    std::size_t (*compressBound)(std::size_t) = nullptr;
    Zstd()
    {
        for (const char *name : { "libzstd.so.1", "libzstd.so" }) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) return;
        createDStream = reinterpret_cast<void *(*)()>(dlsym(lib, "ZSTD_createDStream"));
        freeDStream = reinterpret_cast<std::size_t (*)(void *)>(dlsym(lib, "ZSTD_freeDStream"));
        decompressStream = reinterpret_cast<std::size_t (*)(void *, ZOut *, ZBuf *)>(dlsym(lib, "ZSTD_decompressStream"));
        isError = reinterpret_cast<unsigned (*)(std::size_t)>(dlsym(lib, "ZSTD_isError"));
        compress = reinterpret_cast<std::size_t (*)(void *, std::size_t, const void *, std::size_t, int)>(dlsym(lib, "ZSTD_compress"));
        compressBound = reinterpret_cast<std::size_t (*)(std::size_t)>(dlsym(lib, "ZSTD_compressBound"));
        if (!createDStream || !freeDStream || !decompressStream || !isError || !compress || !compressBound) lib = nullptr;
    }
};
Zstd &zstd()
{
    static Zstd z;
    return z;
}

bytes zstd_decompress(const std::uint8_t *in, std::size_t n)
{
    Zstd &z = zstd();
    if (!z.lib) throw std::logic_error("unsupported compression mode: zstd (libzstd.so.1 not found)");
    void *ds = z.createDStream();
    bytes out;
    std::vector<std::uint8_t> chunk(1 << 17);
    ZBuf ib{ in, n, 0 };
    while (ib.pos < ib.size) {  // concatenated frames are decoded one after the other
        ZOut ob{ chunk.data(), chunk.size(), 0 };
        const std::size_t r = z.decompressStream(ds, &ob, &ib);
        if (z.isError(r)) {
            z.freeDStream(ds);
            throw std::logic_error("stream decompression failed");
        }
        out.insert(out.end(), chunk.begin(), chunk.begin() + (std::ptrdiff_t)ob.pos);
        if (ob.pos == 0 && ib.pos == ib.size) break;
    }
    z.freeDStream(ds);
    return out;
}
bytes zstd_compress(const bytes &in)
{
    Zstd &z = zstd();
    if (!z.lib) throw std::logic_error("unsupported compression mode: zstd (libzstd.so.1 not found)");
    bytes out(z.compressBound(in.size()));
    const std::size_t r = z.compress(out.data(), out.size(), in.data(), in.size(), 3);
    if (z.isError(r)) throw std::logic_error("stream compression failed");
    out.resize(r);
    return out;
}
bytes zlib_decompress(const std::uint8_t *in, std::size_t n)
{
    z_stream s{};
    if (inflateInit(&s) != Z_OK) throw std::logic_error("stream decompression failed");
    s.next_in = const_cast<Bytef *>(in);
    s.avail_in = (uInt)n;
    bytes out;
    std::vector<std::uint8_t> chunk(1 << 17);
    int rc = Z_OK;
    while (rc != Z_STREAM_END) {
        s.next_out = chunk.data();
        s.avail_out = (uInt)chunk.size();
        rc = inflate(&s, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END) {
            inflateEnd(&s);
            throw std::logic_error("stream decompression failed");
        }
        out.insert(out.end(), chunk.begin(), chunk.begin() + (std::ptrdiff_t)(chunk.size() - s.avail_out));
    }
    inflateEnd(&s);
    return out;
}
bytes zlib_compress(const bytes &in)
{
    uLongf cap = compressBound((uLong)in.size());
    bytes out(cap);
    if (compress2(out.data(), &cap, in.data(), (uLong)in.size(), Z_DEFAULT_COMPRESSION) != Z_OK) throw std::logic_error("stream compression failed");
    out.resize(cap);
    return out;
}

// ------------------------------------------------------------------ little helpers over byte strings
struct Reader {
    const std::uint8_t *p;
    std::size_t n, off = 0;
    void need(std::size_t k) const
    {
        if (off + k > n) throw std::logic_error("I/O error: unexpected end of serialized data");
    }
    template <class T>
    T get()
    {
        need(sizeof(T));
        T v;
        std::memcpy(&v, p + off, sizeof(T));
        off += sizeof(T);
        return v;
    }
    void get(void *dst, std::size_t k)
    {
        need(k);
        std::memcpy(dst, p + off, k);
        off += k;
    }
    bytes object()  // a nested container
    {
        std::size_t used = 0;
        bytes m = unwrap(p + off, n - off, &used);
        off += used;
        return m;
    }
};
template <class T>
void put(bytes &b, const T &v)
{
    const std::uint8_t *p = reinterpret_cast<const std::uint8_t *>(&v);
    b.insert(b.end(), p, p + sizeof(T));
}
void put(bytes &b, const void *src, std::size_t k)
{
    const std::uint8_t *p = static_cast<const std::uint8_t *>(src);
    b.insert(b.end(), p, p + k);
}
void append(bytes &b, const bytes &o) { b.insert(b.end(), o.begin(), o.end()); }

constexpr std::uint16_t MAGIC = 0xA15E;
constexpr std::uint8_t HEADER_SIZE = 0x10, VERSION_MAJOR = 4, VERSION_MINOR = 1;

bytes dynarray_members(const std::uint64_t *data, std::uint64_t count)
{
    bytes m;
    put(m, count);
    put(m, data, (std::size_t)count * 8);
    return m;
}
}  // namespace

bool zstd_available() { return zstd().lib != nullptr; }

bytes unwrap(const std::uint8_t *in, std::size_t avail, std::size_t *consumed)
{
    if (avail < HEADER_SIZE) throw std::logic_error("loaded SEALHeader is invalid");
    std::uint16_t magic;
    std::uint64_t size;
    std::memcpy(&magic, in, 2);
    std::memcpy(&size, in + 8, 8);
    const std::uint8_t hs = in[2], vmaj = in[3], mode = in[5];
    if (magic != MAGIC || hs != HEADER_SIZE || size < HEADER_SIZE || size > avail) throw std::logic_error("loaded SEALHeader is invalid");
    if (vmaj != 3 && vmaj != 4) throw std::logic_error("incompatible version");
    if (consumed) *consumed = (std::size_t)size;
    const std::uint8_t *body = in + HEADER_SIZE;
    const std::size_t blen = (std::size_t)size - HEADER_SIZE;
    switch (mode) {
    case 0: return bytes(body, body + blen);
    case 1: return zlib_decompress(body, blen);
    case 2: return zstd_decompress(body, blen);
    default: throw std::logic_error("unsupported compression mode");
    }
}

bytes wrap(const bytes &members, compr_mode mode)
{
    bytes body = mode == compr_mode::none ? members : mode == compr_mode::zlib ? zlib_compress(members) : zstd_compress(members);
    bytes out;
    put(out, MAGIC);
    put(out, HEADER_SIZE);
    put(out, VERSION_MAJOR);
    put(out, VERSION_MINOR);
    put(out, (std::uint8_t)mode);
    put(out, (std::uint16_t)0);
    put(out, (std::uint64_t)(HEADER_SIZE + body.size()));
    append(out, body);
    return out;
}

// ------------------------------------------------------------------ EncryptionParameters
parms_id_t Parms::parms_id(std::size_t limbs) const
{
    if (limbs == 0 || limbs > moduli.size()) throw std::invalid_argument("level out of range");
    std::vector<std::uint64_t> words{ scheme, n };
    words.insert(words.end(), moduli.begin(), moduli.begin() + (std::ptrdiff_t)limbs);
    words.push_back(plain_modulus);
    parms_id_t id;
    blake2b(reinterpret_cast<std::uint8_t *>(id.data()), 32, words.data(), words.size() * 8);
    return id;
}
std::size_t Parms::limbs_of(const parms_id_t &id) const
{
    for (std::size_t l = moduli.size(); l >= 1; --l)
        if (parms_id(l) == id) return l;
    return 0;
}

std::size_t load_parms(const std::uint8_t *in, std::size_t avail, Parms &out)
{
    std::size_t used = 0;
    const bytes m = unwrap(in, avail, &used);
    Reader r{ m.data(), m.size() };
    out.scheme = r.get<std::uint8_t>();
    out.n = r.get<std::uint64_t>();
    const std::uint64_t k = r.get<std::uint64_t>();
    if (k == 0 || k > 64) throw std::logic_error("coeff_modulus is invalid");
    out.moduli.resize((std::size_t)k);
    for (auto &q : out.moduli) {
        const bytes mm = r.object();  // Modulus::save: its own container around the value
        if (mm.size() != 8) throw std::logic_error("Modulus is invalid");
        std::memcpy(&q, mm.data(), 8);
    }
    const bytes pm = r.object();
    if (pm.size() != 8) throw std::logic_error("Modulus is invalid");
    std::memcpy(&out.plain_modulus, pm.data(), 8);
    return used;
}

bytes save_parms(const Parms &p, compr_mode mode)
{
    bytes m;
    put(m, p.scheme);
    put(m, p.n);
    put(m, (std::uint64_t)p.moduli.size());
    auto modulus = [&](std::uint64_t v) {
        bytes mm;
        put(mm, v);
        append(m, wrap(mm, compr_mode::none));
    };
    for (auto q : p.moduli) modulus(q);
    modulus(p.plain_modulus);
    return wrap(m, mode);
}

// ------------------------------------------------------------------ Ciphertext
namespace {
void ct_from_members(const Parms &p, const bytes &m, CtData &c)
{
    Reader r{ m.data(), m.size() };
    r.get(c.parms_id.data(), 32);
    c.is_ntt_form = r.get<std::uint8_t>() != 0;
    c.size = r.get<std::uint64_t>();
    c.n = r.get<std::uint64_t>();
    c.limbs = r.get<std::uint64_t>();
    c.correction_factor = r.get<std::uint64_t>();
    c.scale = r.get<double>();
    if (c.n != p.n || c.limbs == 0 || c.limbs > p.moduli.size() || c.size < 2 || c.size > 6) throw std::logic_error("ciphertext data is invalid");
    if (p.limbs_of(c.parms_id) != c.limbs) throw std::logic_error("ciphertext data is invalid");
    const std::uint64_t total = c.size * c.limbs * c.n;
    const bytes arr = r.object();  // DynArray: u64 count + words
    if (arr.size() < 8) throw std::logic_error("ciphertext data is invalid");
    std::uint64_t count;
    std::memcpy(&count, arr.data(), 8);
    if (arr.size() != 8 + count * 8) throw std::logic_error("ciphertext data is invalid");
    c.data.assign((std::size_t)total, 0);
    if (count == total) {
        std::memcpy(c.data.data(), arr.data() + 8, (std::size_t)total * 8);
        c.was_seeded = false;
    } else if (c.size == 2 && count == total / 2) {  // seeded: polynomial 1 = expansion of the stored seed
        std::memcpy(c.data.data(), arr.data() + 8, (std::size_t)count * 8);
        const bytes info = r.object();  // UniformRandomGeneratorInfo: prng_type + seed
        if (info.size() != 1 + 64) throw std::logic_error("UniformRandomGeneratorInfo is invalid");
        if (info[0] != 1) throw std::logic_error("unsupported prng_type (only blake2xb)");
        std::array<std::uint64_t, 8> seed;
        std::memcpy(seed.data(), info.data() + 1, 64);
        Blake2xbPrng prng(seed);
        const std::vector<std::uint64_t> mods(p.moduli.begin(), p.moduli.begin() + (std::ptrdiff_t)c.limbs);
        sample_poly_uniform(prng, mods, (std::size_t)c.n, c.data.data() + count);
        c.was_seeded = true;
    } else {
        throw std::logic_error("ciphertext data is invalid");
    }
}

bytes ct_members(const CtData &c, const std::array<std::uint64_t, 8> *seed)
{
    bytes m;
    put(m, c.parms_id.data(), 32);
    put(m, (std::uint8_t)(c.is_ntt_form ? 1 : 0));
    put(m, c.size);
    put(m, c.n);
    put(m, c.limbs);
    put(m, c.correction_factor);
    put(m, c.scale);
    const std::uint64_t total = c.size * c.limbs * c.n;
    if (c.data.size() != total) throw std::invalid_argument("ciphertext data has the wrong length");
    if (seed) {
        if (c.size != 2) throw std::invalid_argument("only size-2 ciphertexts can be seeded");
        append(m, wrap(dynarray_members(c.data.data(), total / 2), compr_mode::none));
        bytes info;
        put(info, (std::uint8_t)1);  // prng_type::blake2xb
        put(info, seed->data(), 64);
        append(m, wrap(info, compr_mode::none));
    } else {
        append(m, wrap(dynarray_members(c.data.data(), total), compr_mode::none));
    }
    return m;
}
}  // namespace

std::size_t load_ciphertext(const Parms &p, const std::uint8_t *in, std::size_t avail, CtData &out)
{
    std::size_t used = 0;
    ct_from_members(p, unwrap(in, avail, &used), out);
    return used;
}
bytes save_ciphertext(const CtData &c, compr_mode mode, const std::array<std::uint64_t, 8> *seed) { return wrap(ct_members(c, seed), mode); }

// ------------------------------------------------------------------ KSwitchKeys (RelinKeys / GaloisKeys)
std::size_t load_kswitch_keys(const Parms &p, const std::uint8_t *in, std::size_t avail, KSwitchData &out)
{
    std::size_t used = 0;
    const bytes m = unwrap(in, avail, &used);
    Reader r{ m.data(), m.size() };
    r.get(out.parms_id.data(), 32);
    if (out.parms_id != p.key_parms_id()) throw std::logic_error("KSwitchKeys data is invalid");
    const std::uint64_t dim1 = r.get<std::uint64_t>();
    if (dim1 > 70000) throw std::logic_error("KSwitchKeys data is invalid");
    out.keys.assign((std::size_t)dim1, {});
    for (auto &digits : out.keys) {
        const std::uint64_t dim2 = r.get<std::uint64_t>();
        if (dim2 > p.moduli.size()) throw std::logic_error("KSwitchKeys data is invalid");
        digits.resize((std::size_t)dim2);
        for (auto &k : digits) {
            const bytes pk = r.object();       // PublicKey::save: container around ...
            std::size_t u2 = 0;
            ct_from_members(p, unwrap(pk.data(), pk.size(), &u2), k);  // ... the ciphertext's own container
            if (k.limbs != p.moduli.size() || k.size != 2) throw std::logic_error("KSwitchKeys data is invalid");
        }
    }
    return used;
}

bytes save_kswitch_keys(const KSwitchData &k, compr_mode mode, const std::vector<std::vector<std::array<std::uint64_t, 8>>> *seeds)
{
    bytes m;
    put(m, k.parms_id.data(), 32);
    put(m, (std::uint64_t)k.keys.size());
    for (std::size_t i = 0; i < k.keys.size(); ++i) {
        put(m, (std::uint64_t)k.keys[i].size());
        for (std::size_t j = 0; j < k.keys[i].size(); ++j) {
            const bytes ct = wrap(ct_members(k.keys[i][j], seeds ? &(*seeds)[i][j] : nullptr), compr_mode::none);
            append(m, wrap(ct, compr_mode::none));
        }
    }
    return wrap(m, mode);
}

std::vector<std::uint64_t> KSwitchData::flat(std::size_t index) const
{
    std::vector<std::uint64_t> out;
    for (const auto &d : keys.at(index)) out.insert(out.end(), d.data.begin(), d.data.end());
    return out;
}

}  // namespace he::wire
