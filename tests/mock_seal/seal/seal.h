// Minimal DECLARATION-ONLY stand-in for Microsoft SEAL 4.1's public API (test infrastructure).
// SEAL is absent from this image; this header lets `g++ -fsyntax-only` type-check the code that is written against
// real SEAL -- include/he_gpu_bridge.hpp and tools/seal_golden.cpp -- so that neither carries elided lines.
// Only the members those two files use are declared, with SEAL 4.1's names and signatures (native/src/seal/*.h).
// Nothing here is ever linked or executed.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace seal {

using parms_id_type = std::array<std::uint64_t, 4>;
enum class scheme_type : std::uint8_t { none = 0, bfv = 1, ckks = 2, bgv = 3 };
enum class sec_level_type : int { none = 0, tc128 = 128, tc192 = 192, tc256 = 256 };

class Modulus {
public:
    std::uint64_t value() const noexcept;
    int bit_count() const noexcept;
};

class CoeffModulus {
public:
    static std::vector<Modulus> Create(std::size_t poly_modulus_degree, std::vector<int> bit_sizes);
};

class EncryptionParameters {
public:
    EncryptionParameters(scheme_type scheme);
    void set_poly_modulus_degree(std::size_t poly_modulus_degree);
    void set_coeff_modulus(const std::vector<Modulus> &coeff_modulus);
    std::size_t poly_modulus_degree() const noexcept;
    const std::vector<Modulus> &coeff_modulus() const noexcept;
};

class SEALContext {
public:
    class ContextData {
    public:
        const EncryptionParameters &parms() const noexcept;
        const parms_id_type &parms_id() const noexcept;
        std::shared_ptr<const ContextData> next_context_data() const noexcept;
        std::size_t chain_index() const noexcept;
    };
    SEALContext(const EncryptionParameters &parms, bool expand_mod_chain = true, sec_level_type sec_level = sec_level_type::tc128);
    std::shared_ptr<const ContextData> get_context_data(parms_id_type parms_id) const;
    std::shared_ptr<const ContextData> key_context_data() const;
    std::shared_ptr<const ContextData> first_context_data() const;
    const parms_id_type &key_parms_id() const noexcept;
    const parms_id_type &first_parms_id() const noexcept;
};

class Plaintext {
public:
    Plaintext();
    void resize(std::size_t coeff_count);
    std::uint64_t *data();
    const std::uint64_t *data() const;
    std::size_t coeff_count() const noexcept;
    parms_id_type &parms_id() noexcept;
    const parms_id_type &parms_id() const noexcept;
    double &scale() noexcept;
    const double &scale() const noexcept;
};

class Ciphertext {
public:
    Ciphertext();
    void resize(const SEALContext &context, parms_id_type parms_id, std::size_t size);
    std::uint64_t *data() noexcept;
    const std::uint64_t *data() const noexcept;
    std::uint64_t *data(std::size_t poly_index);
    std::size_t size() const noexcept;
    std::size_t coeff_modulus_size() const noexcept;
    std::size_t poly_modulus_degree() const noexcept;
    bool &is_ntt_form() noexcept;
    const bool &is_ntt_form() const noexcept;
    parms_id_type &parms_id() noexcept;
    const parms_id_type &parms_id() const noexcept;
    double &scale() noexcept;
    const double &scale() const noexcept;
};

class PublicKey {
public:
    Ciphertext &data() noexcept;
    const Ciphertext &data() const noexcept;
};

class SecretKey {
public:
    Plaintext &data() noexcept;
    const Plaintext &data() const noexcept;
};

class KSwitchKeys {
public:
    std::vector<std::vector<PublicKey>> &data() noexcept;
    const std::vector<std::vector<PublicKey>> &data() const noexcept;
};

class RelinKeys : public KSwitchKeys {
public:
    static std::size_t get_index(std::size_t key_power);
};

class GaloisKeys : public KSwitchKeys {
public:
    static std::size_t get_index(std::uint32_t galois_elt);
    bool has_key(std::uint32_t galois_elt) const;
};

class KeyGenerator {
public:
    KeyGenerator(const SEALContext &context);
    const SecretKey &secret_key() const;
    void create_public_key(PublicKey &destination) const;
    void create_relin_keys(RelinKeys &destination);
    void create_galois_keys(GaloisKeys &destination);
    void create_galois_keys(const std::vector<int> &steps, GaloisKeys &destination);
};

class Encryptor {
public:
    Encryptor(const SEALContext &context, const PublicKey &public_key);
    void encrypt(const Plaintext &plain, Ciphertext &destination) const;
};

class CKKSEncoder {
public:
    CKKSEncoder(const SEALContext &context);
    void encode(const std::vector<double> &values, double scale, Plaintext &destination);
    void encode(const std::vector<double> &values, parms_id_type parms_id, double scale, Plaintext &destination);
    std::size_t slot_count() const noexcept;
};

class Evaluator {
public:
    Evaluator(const SEALContext &context);
    void negate(const Ciphertext &encrypted, Ciphertext &destination) const;
    void add(const Ciphertext &encrypted1, const Ciphertext &encrypted2, Ciphertext &destination) const;
    void sub(const Ciphertext &encrypted1, const Ciphertext &encrypted2, Ciphertext &destination) const;
    void multiply(const Ciphertext &encrypted1, const Ciphertext &encrypted2, Ciphertext &destination) const;
    void square(const Ciphertext &encrypted, Ciphertext &destination) const;
    void relinearize(const Ciphertext &encrypted, const RelinKeys &relin_keys, Ciphertext &destination) const;
    void rescale_to_next(const Ciphertext &encrypted, Ciphertext &destination) const;
    void mod_switch_to_next(const Ciphertext &encrypted, Ciphertext &destination) const;
    void add_plain(const Ciphertext &encrypted, const Plaintext &plain, Ciphertext &destination) const;
    void sub_plain(const Ciphertext &encrypted, const Plaintext &plain, Ciphertext &destination) const;
    void multiply_plain(const Ciphertext &encrypted, const Plaintext &plain, Ciphertext &destination) const;
    void rotate_vector(const Ciphertext &encrypted, int steps, const GaloisKeys &galois_keys, Ciphertext &destination) const;
};

}  // namespace seal
