// seal_wire.hpp -- Microsoft SEAL 4.1's wire format for the objects the reference's client / server exchange
// (src/demos/client.cpp:66-319, src/demos/server.cpp:92-244,527-592): EncryptionParameters,
// Serializable<RelinKeys> (seeded key-switching keys), seeded symmetric ciphertexts in, plain ciphertexts out.
//
// SEAL is absent from this image and the reference holds no serialized bytes, so this is a restatement of SEAL
// 4.1's published sources (native/src/seal/serialization.h, encryptionparams.cpp, ciphertext.cpp, kswitchkeys.cpp,
// dynarray.h, randomgen.cpp, util/rlwe.cpp, util/blake2xb.c) -- **format fidelity unpinned against real SEAL**:
// BLAKE2b is checked against RFC 7693 / hashlib, the container, the seed expansion and the sampler against an
// independent Python restatement (tests/seal_wire_ref.py).  Pure host code: nothing here touches the GPU.
//
//   container   16-byte header {u16 0xA15E, u8 0x10, u8 major, u8 minor, u8 compr_mode, u16 0, u64 total size}
//               followed by the members, raw (compr_mode 0), zlib (1) or zstd (2) compressed
//   members     nested objects (Modulus, DynArray, UniformRandomGeneratorInfo, PublicKey, Ciphertext) carry their
//               own uncompressed container inside the outer payload
//   seeds       a seeded ciphertext stores polynomial 0 and {prng_type, 64-byte seed}; polynomial 1 is
//               sample_poly_uniform over the BLAKE2xb PRNG (4096-byte buffers, buffer k = BLAKE2xb(counter k, key = seed))
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace he::wire {

using bytes = std::vector<std::uint8_t>;
using parms_id_t = std::array<std::uint64_t, 4>;

// ---- BLAKE2b (RFC 7693) with the full parameter block, and BLAKE2xb (the XOF SEAL's default PRNG uses)
void blake2b(std::uint8_t *out, std::size_t outlen, const void *in, std::size_t inlen, const void *key = nullptr, std::size_t keylen = 0);
void blake2b_param(std::uint8_t *out, std::size_t outlen, const std::uint8_t param[64], const void *in, std::size_t inlen, const void *key,
                   std::size_t keylen);
void blake2xb(std::uint8_t *out, std::size_t outlen, const void *in, std::size_t inlen, const void *key, std::size_t keylen);

// seal::Blake2xbPRNG (randomgen.cpp): 4096-byte buffer, refill k = blake2xb(buffer, 4096, &counter_k, 8, seed, 64)
class Blake2xbPrng {
public:
    explicit Blake2xbPrng(const std::array<std::uint64_t, 8> &seed) : seed_(seed) {}
    void generate(std::size_t count, std::uint8_t *dst);

private:
    std::array<std::uint64_t, 8> seed_;
    std::uint64_t counter_ = 0;
    std::array<std::uint8_t, 4096> buf_{};
    std::size_t pos_ = 4096;
};
// seal::util::sample_poly_uniform (util/rlwe.cpp): fill [limbs][n] with PRNG bytes, then per word rejection-sample
// against the largest multiple of the modulus below 2^64 and reduce
void sample_poly_uniform(Blake2xbPrng &prng, const std::vector<std::uint64_t> &moduli, std::size_t n, std::uint64_t *dst);

// ---- container
enum class compr_mode : std::uint8_t { none = 0, zlib = 1, zstd = 2 };
bool zstd_available();
// members of the object that starts at `in`; *consumed = its total size (header included)
bytes unwrap(const std::uint8_t *in, std::size_t avail, std::size_t *consumed);
bytes wrap(const bytes &members, compr_mode mode = compr_mode::none);

// ---- objects (raw limb data in SEAL's layouts)
struct Parms {
    std::uint8_t scheme = 2;  // scheme_type::ckks
    std::uint64_t n = 0;
    std::vector<std::uint64_t> moduli;  // key level: data primes + special prime
    std::uint64_t plain_modulus = 0;
    parms_id_t parms_id(std::size_t limbs) const;  // BLAKE2b-256 of [scheme, n, q_0..q_{limbs-1}, plain_modulus]
    parms_id_t key_parms_id() const { return parms_id(moduli.size()); }
    std::size_t limbs_of(const parms_id_t &id) const;  // 0 if the id is not a level of this chain
};
struct CtData {
    parms_id_t parms_id{};
    bool is_ntt_form = true;
    std::uint64_t size = 0, n = 0, limbs = 0, correction_factor = 1;
    double scale = 1.0;
    std::vector<std::uint64_t> data;  // [size][limbs][n]
    bool was_seeded = false;          // load: polynomial 1 came from a seed
};
struct KSwitchData {
    parms_id_t parms_id{};
    std::vector<std::vector<CtData>> keys;  // [index][digit] -> size-2 ciphertext over the key-level limbs
    // digits of one index flattened to the C ABI's key layout [digits][2][K][n]
    std::vector<std::uint64_t> flat(std::size_t index) const;
};

std::size_t load_parms(const std::uint8_t *in, std::size_t avail, Parms &out);
bytes save_parms(const Parms &p, compr_mode mode = compr_mode::none);
std::size_t load_ciphertext(const Parms &p, const std::uint8_t *in, std::size_t avail, CtData &out);
// seed != nullptr: the seeded form (polynomial 1 replaced by the seed; size must be 2 and data(1) must be the seed's expansion)
bytes save_ciphertext(const CtData &c, compr_mode mode = compr_mode::none, const std::array<std::uint64_t, 8> *seed = nullptr);
std::size_t load_kswitch_keys(const Parms &p, const std::uint8_t *in, std::size_t avail, KSwitchData &out);
bytes save_kswitch_keys(const KSwitchData &k, compr_mode mode = compr_mode::none, const std::vector<std::vector<std::array<std::uint64_t, 8>>> *seeds = nullptr);

}  // namespace he::wire
