// he_gpu_bridge.hpp -- the reference-side binding: seal:: objects <-> libhegpu.so (C ABI, include/hegpu.h).
//
// Header only; compiled by the reference's maintainer where Microsoft SEAL 4.1 is installed.
// This image has no SEAL, so the repo type-checks it against tests/mock_seal/seal/seal.h (declarations of the
// SEAL 4.1 members used here; tests/test_abi.py::test_seal_facing_sources_compile) -- no elided lines.
//
// It replaces, at the reference's own call sites, the seal::Evaluator calls of src/core/he_operators.cpp:14-237,
// src/core/he_linalg.cpp:595,609,622,636,647 and include/he_util.h:35-36,43-44: same argument meaning, same
// exception types (std::invalid_argument / std::logic_error carrying SEAL's messages).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "hegpu.h"
#include "seal/seal.h"

namespace he::gpu {

inline void check(int st)  // status -> SEAL's exception types
{
    if (st == HEGPU_OK) return;
    if (st == HEGPU_ERR_INVALID_ARGUMENT) throw std::invalid_argument(hegpu_last_error());
    if (st == HEGPU_ERR_LOGIC) throw std::logic_error(hegpu_last_error());
    throw std::runtime_error(hegpu_last_error());
}

// RAII handles of device-resident batches
struct DeviceCt {
    hegpu_ct *h = nullptr;
    DeviceCt() = default;
    DeviceCt(const DeviceCt &) = delete;
    DeviceCt &operator=(const DeviceCt &) = delete;
    DeviceCt(DeviceCt &&o) noexcept : h(o.h) { o.h = nullptr; }
    ~DeviceCt() { if (h) hegpu_ct_destroy(h); }
};
struct DevicePt {
    hegpu_pt *h = nullptr;
    DevicePt() = default;
    DevicePt(const DevicePt &) = delete;
    DevicePt &operator=(const DevicePt &) = delete;
    DevicePt(DevicePt &&o) noexcept : h(o.h) { o.h = nullptr; }
    ~DevicePt() { if (h) hegpu_pt_destroy(h); }
};

// one per seal::SEALContext; built from the key-level parameters
class Bridge {
public:
    hegpu_ctx *ctx = nullptr;
    const seal::SEALContext &seal_ctx;
    std::size_t n = 0, K = 0;  // ring degree, primes of the key level (last = special prime)

    explicit Bridge(const seal::SEALContext &c, int device = 0) : seal_ctx(c)
    {
        const auto &parms = c.key_context_data()->parms();
        std::vector<std::uint64_t> q;
        for (const auto &m : parms.coeff_modulus()) q.push_back(m.value());  // CoeffModulus::Create order
        n = parms.poly_modulus_degree();
        K = q.size();
        check(hegpu_ctx_create(&ctx, (std::uint32_t)n, q.data(), (std::uint32_t)K, device));
    }
    Bridge(const Bridge &) = delete;
    Bridge &operator=(const Bridge &) = delete;
    ~Bridge() { hegpu_ctx_destroy(ctx); }

    // level of a SEAL object = number of limbs of its parms_id
    std::uint32_t limbs(const seal::parms_id_type &id) const
    {
        return (std::uint32_t)seal_ctx.get_context_data(id)->parms().coeff_modulus().size();
    }
    seal::parms_id_type parms_at(std::uint32_t L) const  // walk the chain down to L limbs
    {
        auto cd = seal_ctx.first_context_data();
        while (cd->parms().coeff_modulus().size() > L) cd = cd->next_context_data();
        return cd->parms_id();
    }

    // keys: KSwitchKeys::data()[idx][j] is a size-2 ciphertext over the K key-level limbs -> [K-1][2][K][N]
    void load(const seal::RelinKeys &rk) { load_key(rk.data()[seal::RelinKeys::get_index(2)], 0, true); }
    void load(const seal::GaloisKeys &gk)
    {
        for (std::size_t i = 0; i < gk.data().size(); ++i)
            if (!gk.data()[i].empty()) load_key(gk.data()[i], (std::uint32_t)(2 * i + 1), false);
    }
    void load_key(const std::vector<seal::PublicKey> &digits, std::uint32_t elt, bool relin)
    {
        const std::size_t words = 2 * K * n;
        std::vector<std::uint64_t> flat(digits.size() * words);
        for (std::size_t j = 0; j < digits.size(); ++j)
            std::memcpy(&flat[j * words], digits[j].data().data(), words * sizeof(std::uint64_t));
        check(relin ? hegpu_load_relin_key(ctx, flat.data()) : hegpu_load_galois_key(ctx, elt, flat.data()));
    }

    // ciphertext: seal::Ciphertext::data() is [size][L][N] contiguous -- the ABI's host layout
    DeviceCt up(const seal::Ciphertext &c) const
    {
        DeviceCt d;
        check(hegpu_ct_create(ctx, &d.h, 1, 3, (std::uint32_t)(K - 1)));
        check(hegpu_ct_upload(d.h, c.data(), (std::uint32_t)c.size(), limbs(c.parms_id()), c.scale()));
        return d;
    }
    DeviceCt up(const std::vector<seal::Ciphertext> &cs) const  // a batch of equally shaped ciphertexts
    {
        DeviceCt d;
        if (cs.empty()) throw std::invalid_argument("empty batch");
        const std::uint32_t size = (std::uint32_t)cs[0].size(), L = limbs(cs[0].parms_id());
        check(hegpu_ct_create(ctx, &d.h, (std::uint32_t)cs.size(), 3, (std::uint32_t)(K - 1)));
        check(hegpu_ct_set_meta(d.h, size, L, cs[0].scale()));
        for (std::size_t i = 0; i < cs.size(); ++i) {
            if (cs[i].size() != size || limbs(cs[i].parms_id()) != L) throw std::invalid_argument("encrypted parameter mismatch");
            check(hegpu_ct_upload_one(d.h, (std::uint32_t)i, cs[i].data()));
        }
        return d;
    }
    DevicePt up(const seal::Plaintext &p) const  // CKKS plaintexts are NTT-form [L][N]
    {
        DevicePt d;
        check(hegpu_pt_create(ctx, &d.h, 1, (std::uint32_t)(K - 1)));
        check(hegpu_pt_upload(d.h, p.data(), limbs(p.parms_id()), p.scale()));
        return d;
    }
    void down(const DeviceCt &d, seal::Ciphertext &dst, std::uint32_t index = 0) const
    {
        std::uint32_t B, size, L;
        double scale;
        check(hegpu_ct_info(d.h, &B, &size, &L, &scale));
        dst.resize(seal_ctx, parms_at(L), size);
        dst.is_ntt_form() = true;  // metadata fidelity
        dst.scale() = scale;
        check(hegpu_ct_download_one(d.h, index, dst.data()));
    }
};

// ---- the operator bodies of src/core/he_operators.cpp re-pointed at the device, one C-ABI call each.
// (The literal per-operator form moves every ciphertext over PCIe twice per operator; the intended integration
// overrides the coarse routines instead -- INTEGRATION.md 1, table.)
inline void negate(Bridge &B, seal::Ciphertext &op)
{
    auto a = B.up(op);
    check(hegpu_negate(B.ctx, a.h, a.h));
    B.down(a, op);
}
inline void add(Bridge &B, seal::Ciphertext &op1, const seal::Ciphertext &op2)
{
    auto a = B.up(op1), b = B.up(op2);
    check(hegpu_add(B.ctx, a.h, a.h, b.h));
    B.down(a, op1);
}
inline void sub(Bridge &B, seal::Ciphertext &op1, const seal::Ciphertext &op2)
{
    auto a = B.up(op1), b = B.up(op2);
    check(hegpu_sub(B.ctx, a.h, a.h, b.h));
    B.down(a, op1);
}
inline void multiply(Bridge &B, seal::Ciphertext &op1, const seal::Ciphertext &op2)
{
    auto a = B.up(op1), b = B.up(op2);
    check(hegpu_multiply(B.ctx, a.h, a.h, b.h));
    B.down(a, op1);
}
inline void add_plain(Bridge &B, seal::Ciphertext &op1, const seal::Plaintext &op2)
{
    auto a = B.up(op1);
    auto p = B.up(op2);
    check(hegpu_add_plain(B.ctx, a.h, a.h, p.h, 0));
    B.down(a, op1);
}
inline void sub_plain(Bridge &B, seal::Ciphertext &op1, const seal::Plaintext &op2)
{
    auto a = B.up(op1);
    auto p = B.up(op2);
    check(hegpu_sub_plain(B.ctx, a.h, a.h, p.h, 0));
    B.down(a, op1);
}
inline void multiply_plain(Bridge &B, seal::Ciphertext &op1, const seal::Plaintext &op2)
{
    auto a = B.up(op1);
    auto p = B.up(op2);
    check(hegpu_multiply_plain(B.ctx, a.h, a.h, p.h, 0));  // replaces eval.multiply_plain_inplace (he_operators.cpp:130,140)
    B.down(a, op1);
}
inline void relinearize(Bridge &B, seal::Ciphertext &op)  // keys loaded once with Bridge::load
{
    auto a = B.up(op);
    check(hegpu_relinearize(B.ctx, a.h, a.h));
    B.down(a, op);
}
inline void rescale_to_next(Bridge &B, seal::Ciphertext &op)
{
    auto a = B.up(op);
    check(hegpu_rescale_to_next(B.ctx, a.h, a.h));
    B.down(a, op);
}
inline void mod_switch_to_next(Bridge &B, seal::Ciphertext &op)
{
    auto a = B.up(op);
    check(hegpu_mod_switch_to_next(B.ctx, a.h, a.h));
    B.down(a, op);
}
inline void rotate_vector(Bridge &B, seal::Ciphertext &op, int steps)  // NAF chains as SEAL's rotate_internal
{
    auto a = B.up(op);
    check(hegpu_rotate_vector(B.ctx, a.h, a.h, steps));
    B.down(a, op);
}

// ---- a coarse override: BatchedMatrix::matmul (src/core/he_linalg.cpp:943-1006), data on the device for the whole
// routine.  this_cts / other_cts are the BatchedVector ciphertexts in the reference's order; result: p ciphertexts.
inline void bmatmul(Bridge &B, const std::vector<seal::Ciphertext> &this_cts, const std::vector<seal::Ciphertext> &other_cts,
                    std::uint32_t n_dim, std::uint32_t p, bool case_b, std::vector<seal::Ciphertext> &result)
{
    auto a = B.up(this_cts), b = B.up(other_cts);
    DeviceCt out;
    check(hegpu_ct_create(B.ctx, &out.h, p, 3, (std::uint32_t)(B.K - 1)));
    check(hegpu_bmatmul(B.ctx, out.h, a.h, b.h, n_dim, p, case_b ? 1 : 0));
    result.resize(p);
    for (std::uint32_t i = 0; i < p; ++i) B.down(out, result[i], i);
}

}  // namespace he::gpu
