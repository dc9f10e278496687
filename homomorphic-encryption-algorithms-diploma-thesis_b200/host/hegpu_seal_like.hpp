// hegpu_seal_like.hpp -- SEAL-shaped host value types over the C ABI (include/hegpu.h).
//
// The reference passes `const seal::Evaluator &` plus seal::Ciphertext / Plaintext /
// RelinKeys / GaloisKeys / CKKSEncoder into every routine (include/he_operators.h:44-159,
// include/he_linalg.h:47-412, include/he_fft.h:12-27).  SEAL is not available in this image,
// so the host mirror of those routines (he_operators / he_linalg / he_fft / he_util in this
// directory) is written against the types below, which keep SEAL's method names, argument
// meaning and exception types but hold their data in HBM (a batch-of-one hegpu_ct) instead
// of host memory: an operator costs kernel launches, not PCIe round trips.
// With real SEAL the same mirror is bound through the bridge shown in INTEGRATION.md.
#pragma once

#include <complex>
#include <cstdlib>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "hegpu.h"

namespace he::gpu {

inline void check(int status)
{
    if (status == HEGPU_OK) return;
    const std::string msg = hegpu_last_error();
    if (status == HEGPU_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    if (status == HEGPU_ERR_LOGIC) throw std::logic_error(msg);
    if (status == HEGPU_ERR_OUT_OF_MEMORY) throw std::bad_alloc();
    throw std::runtime_error(msg);
}

// seal::parms_id_type stand-in: a level of the modulus chain = its number of RNS limbs
struct parms_id_type {
    std::uint32_t limbs = 0;
    bool operator==(const parms_id_type &o) const { return limbs == o.limbs; }
    bool operator!=(const parms_id_type &o) const { return limbs != o.limbs; }
};

// seal::SEALContext stand-in (owns the hegpu_ctx)
class SEALContext {
public:
    SEALContext(std::uint32_t poly_modulus_degree, const std::vector<std::uint64_t> &coeff_modulus, int device = 0)
        : n_(poly_modulus_degree), moduli_(coeff_modulus)
    {
        hegpu_ctx *c = nullptr;
        check(hegpu_ctx_create(&c, n_, moduli_.data(), (std::uint32_t)moduli_.size(), device));
        ctx_.reset(c, [](hegpu_ctx *p) { hegpu_ctx_destroy(p); });
    }
    hegpu_ctx *raw() const { return ctx_.get(); }
    std::uint32_t poly_modulus_degree() const { return n_; }
    const std::vector<std::uint64_t> &coeff_modulus() const { return moduli_; }
    std::uint32_t key_limbs() const { return (std::uint32_t)moduli_.size(); }
    parms_id_type first_parms_id() const { return parms_id_type{ key_limbs() - 1 }; }
    // chain_index counts down to 0 at the last level (he_util.h:13-16)
    std::size_t chain_index(const parms_id_type &id) const { return id.limbs - 1; }
    void sync() const { check(hegpu_sync(raw())); }

private:
    std::uint32_t n_;
    std::vector<std::uint64_t> moduli_;
    std::shared_ptr<hegpu_ctx> ctx_;
};

// seal::Ciphertext stand-in: value semantics (deep copies, like he_linalg.h:51-55 relies on)
class Ciphertext {
public:
    Ciphertext() = default;
    explicit Ciphertext(const SEALContext &ctx) : ctx_(&ctx) {}
    Ciphertext(const Ciphertext &o) : ctx_(o.ctx_)
    {
        if (o.h_) {
            ensure();
            check(hegpu_ct_copy(ctx_->raw(), h_.get(), o.h_.get()));
        }
    }
    Ciphertext(Ciphertext &&) = default;
    Ciphertext &operator=(const Ciphertext &o)
    {
        if (this != &o) {
            Ciphertext tmp(o);
            *this = std::move(tmp);
        }
        return *this;
    }
    Ciphertext &operator=(Ciphertext &&) = default;

    // host [size][L][N] (seal::Ciphertext::data() layout)
    void load(const SEALContext &ctx, const std::uint64_t *host, std::uint32_t size, std::uint32_t limbs, double scale)
    {
        ctx_ = &ctx;
        ensure();
        check(hegpu_ct_upload(h_.get(), host, size, limbs, scale));
    }
    std::vector<std::uint64_t> save() const
    {
        std::vector<std::uint64_t> out((std::size_t)size() * coeff_modulus_size() * ctx_->poly_modulus_degree());
        check(hegpu_ct_download(h_.get(), out.data()));
        return out;
    }
    std::size_t size() const { return info().size; }
    std::size_t coeff_modulus_size() const { return info().limbs; }
    parms_id_type parms_id() const { return parms_id_type{ info().limbs }; }
    double scale() const { return info().scale; }
    void set_scale(double s) { check(hegpu_ct_set_scale(handle(), s)); }
    bool is_ntt_form() const { return true; }

    const SEALContext *context() const { return ctx_; }
    hegpu_ct *handle() const
    {
        if (!h_) throw std::invalid_argument("encrypted is not valid for encryption parameters");
        return h_.get();
    }
    // destination of an out-of-place evaluator call
    hegpu_ct *prepare(const SEALContext &ctx)
    {
        ctx_ = &ctx;
        ensure();
        return h_.get();
    }

private:
    struct Info {
        std::uint32_t size, limbs;
        double scale;
    };
    Info info() const
    {
        Info i{ 0, 0, 1.0 };
        std::uint32_t b;
        if (h_) check(hegpu_ct_info(h_.get(), &b, &i.size, &i.limbs, &i.scale));
        return i;
    }
    void ensure()
    {
        if (h_) return;
        hegpu_ct *p = nullptr;
        check(hegpu_ct_create(ctx_->raw(), &p, 1, 3, ctx_->key_limbs() - 1));
        h_.reset(p, [](hegpu_ct *q) { hegpu_ct_destroy(q); });
    }
    const SEALContext *ctx_ = nullptr;
    std::shared_ptr<hegpu_ct> h_;  // unique per object; shared_ptr only for the custom deleter + move
};

// seal::Plaintext stand-in (NTT form at a level)
class Plaintext {
public:
    Plaintext() = default;
    void load(const SEALContext &ctx, const std::uint64_t *host, std::uint32_t limbs, double scale)
    {
        hegpu_pt *p = nullptr;
        check(hegpu_pt_create(ctx.raw(), &p, 1, limbs));
        h_.reset(p, [](hegpu_pt *q) { hegpu_pt_destroy(q); });
        check(hegpu_pt_upload(p, host, limbs, scale));
        limbs_ = limbs;
        scale_ = scale;
        n_ = ctx.poly_modulus_degree();
    }
    std::vector<std::uint64_t> save() const
    {
        std::vector<std::uint64_t> out((std::size_t)limbs_ * n_);
        check(hegpu_pt_download_one(handle(), 0, out.data()));
        return out;
    }
    hegpu_pt *handle() const
    {
        if (!h_) throw std::invalid_argument("plain is not valid for encryption parameters");
        return h_.get();
    }
    parms_id_type parms_id() const { return parms_id_type{ limbs_ }; }
    double scale() const { return scale_; }

private:
    std::shared_ptr<hegpu_pt> h_;
    std::uint32_t limbs_ = 0, n_ = 0;
    double scale_ = 1.0;
};

// seal::RelinKeys / seal::GaloisKeys stand-ins: the key material lives in the context
// (hegpu_load_relin_key / hegpu_load_galois_key); these are the tokens the reference's
// signatures pass around.
class RelinKeys {
public:
    RelinKeys() = default;
    // host [Lmax][2][K][N] == KSwitchKeys::data()[0]
    void load(const SEALContext &ctx, const std::uint64_t *host) { check(hegpu_load_relin_key(ctx.raw(), host)); }
};
class GaloisKeys {
public:
    GaloisKeys() = default;
    void load(const SEALContext &ctx, std::uint32_t galois_elt, const std::uint64_t *host)
    {
        check(hegpu_load_galois_key(ctx.raw(), galois_elt, host));
    }
    bool has_key(const SEALContext &ctx, std::uint32_t galois_elt) const { return hegpu_has_galois_key(ctx.raw(), galois_elt) != 0; }
};

// seal::Evaluator stand-in: the 17 methods of SURVEY 8b, same names and argument order
class Evaluator {
public:
    explicit Evaluator(const SEALContext &ctx) : ctx_(ctx)
    {
        if (const char *e = std::getenv("HEGPU_THROW_ON_TRANSPARENT")) throw_on_transparent = std::atoi(e) != 0;
    }
    const SEALContext &context() const { return ctx_; }
    // SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT (SEAL's default build): every operation ends with
    // `if (result.is_transparent()) throw std::logic_error("result ciphertext is transparent")`.  Here the check is
    // a device reduction plus a blocking read, so it is opt-in (this flag or HEGPU_THROW_ON_TRANSPARENT=1).
    bool throw_on_transparent = false;

    void negate_inplace(Ciphertext &a) const { check(hegpu_negate(c(), a.handle(), a.handle())); post(a); }
    void negate(const Ciphertext &a, Ciphertext &d) const { check(hegpu_negate(c(), d.prepare(ctx_), a.handle())); post(d); }
    void add_inplace(Ciphertext &a, const Ciphertext &b) const { check(hegpu_add(c(), a.handle(), a.handle(), b.handle())); post(a); }
    void add(const Ciphertext &a, const Ciphertext &b, Ciphertext &d) const { check(hegpu_add(c(), d.prepare(ctx_), a.handle(), b.handle())); post(d); }
    void sub_inplace(Ciphertext &a, const Ciphertext &b) const { check(hegpu_sub(c(), a.handle(), a.handle(), b.handle())); post(a); }
    void sub(const Ciphertext &a, const Ciphertext &b, Ciphertext &d) const { check(hegpu_sub(c(), d.prepare(ctx_), a.handle(), b.handle())); post(d); }
    void add_plain_inplace(Ciphertext &a, const Plaintext &p) const { check(hegpu_add_plain(c(), a.handle(), a.handle(), p.handle(), 0)); post(a); }
    void add_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &d) const { check(hegpu_add_plain(c(), d.prepare(ctx_), a.handle(), p.handle(), 0)); post(d); }
    void sub_plain_inplace(Ciphertext &a, const Plaintext &p) const { check(hegpu_sub_plain(c(), a.handle(), a.handle(), p.handle(), 0)); post(a); }
    void sub_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &d) const { check(hegpu_sub_plain(c(), d.prepare(ctx_), a.handle(), p.handle(), 0)); post(d); }
    void multiply_inplace(Ciphertext &a, const Ciphertext &b) const { check(hegpu_multiply(c(), a.handle(), a.handle(), b.handle())); post(a); }
    void multiply(const Ciphertext &a, const Ciphertext &b, Ciphertext &d) const { check(hegpu_multiply(c(), d.prepare(ctx_), a.handle(), b.handle())); post(d); }
    void square_inplace(Ciphertext &a) const { check(hegpu_square(c(), a.handle(), a.handle())); post(a); }
    void square(const Ciphertext &a, Ciphertext &d) const { check(hegpu_square(c(), d.prepare(ctx_), a.handle())); post(d); }
    void multiply_plain_inplace(Ciphertext &a, const Plaintext &p) const { check(hegpu_multiply_plain(c(), a.handle(), a.handle(), p.handle(), 0)); post(a); }
    void multiply_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &d) const { check(hegpu_multiply_plain(c(), d.prepare(ctx_), a.handle(), p.handle(), 0)); post(d); }
    void relinearize_inplace(Ciphertext &a, const RelinKeys &) const { check(hegpu_relinearize(c(), a.handle(), a.handle())); post(a); }
    void relinearize(const Ciphertext &a, const RelinKeys &, Ciphertext &d) const { check(hegpu_relinearize(c(), d.prepare(ctx_), a.handle())); post(d); }
    void rescale_to_next_inplace(Ciphertext &a) const { check(hegpu_rescale_to_next(c(), a.handle(), a.handle())); post(a); }
    void rescale_to_next(const Ciphertext &a, Ciphertext &d) const { check(hegpu_rescale_to_next(c(), d.prepare(ctx_), a.handle())); post(d); }
    void mod_switch_to_next_inplace(Ciphertext &a) const { check(hegpu_mod_switch_to_next(c(), a.handle(), a.handle())); post(a); }
    void mod_switch_to_next(const Ciphertext &a, Ciphertext &d) const { check(hegpu_mod_switch_to_next(c(), d.prepare(ctx_), a.handle())); post(d); }
    void rotate_vector_inplace(Ciphertext &a, int steps, const GaloisKeys &) const { check(hegpu_rotate_vector(c(), a.handle(), a.handle(), steps)); post(a); }
    void rotate_vector(const Ciphertext &a, int steps, const GaloisKeys &, Ciphertext &d) const
    {
        check(hegpu_rotate_vector(c(), d.prepare(ctx_), a.handle(), steps));
        post(d);
    }

private:
    hegpu_ctx *c() const { return ctx_.raw(); }
    void post(const Ciphertext &r) const
    {
        if (!throw_on_transparent) return;
        std::uint32_t cnt = 0;
        check(hegpu_ct_transparent(c(), r.handle(), &cnt));
        if (cnt) throw std::logic_error("result ciphertext is transparent");
    }
    const SEALContext &ctx_;
};

// seal::CKKSEncoder stand-in.  The double-precision embedding FFT runs on the
// host exactly as in SEAL; the RNS NTTs of the result run on the GPU.
class CKKSEncoder {
public:
    explicit CKKSEncoder(const SEALContext &ctx);
    std::size_t slot_count() const { return slots_; }
    void encode(const std::vector<std::complex<double>> &values, parms_id_type parms_id, double scale, Plaintext &dst) const;
    void encode(const std::vector<double> &values, parms_id_type parms_id, double scale, Plaintext &dst) const;
    void encode(std::complex<double> value, parms_id_type parms_id, double scale, Plaintext &dst) const
    {
        encode(std::vector<std::complex<double>>(slots_, value), parms_id, scale, dst);
    }
    // SEAL's real-scalar shortcut: round(value*scale) in every NTT slot
    void encode(double value, parms_id_type parms_id, double scale, Plaintext &dst) const;

private:
    const SEALContext &ctx_;
    std::size_t n_, slots_;
    std::vector<std::uint32_t> idx1_, idx2_;
    std::vector<std::complex<double>> zeta_neg_;
};

}  // namespace he::gpu
