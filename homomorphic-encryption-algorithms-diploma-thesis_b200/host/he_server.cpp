// he_server.cpp -- see he_server.hpp.
#include "he_server.hpp"

#include "he_fft.h"
#include "he_linalg.h"
#include "he_operators.h"

namespace he::server {

using namespace he::gpu;
using namespace he::operators;
using he::wire::bytes;

namespace {
struct Session {
    he::wire::Parms parms;
    std::size_t pos = 0;
    const bytes &buf;
    explicit Session(const bytes &b) : buf(b) { pos += he::wire::load_parms(buf.data(), buf.size(), parms); }
    SEALContext context(int device) const { return SEALContext((std::uint32_t)parms.n, parms.moduli, device); }
    void load_relin_keys(const SEALContext &ctx, RelinKeys &rk)
    {
        he::wire::KSwitchData k;
        pos += he::wire::load_kswitch_keys(parms, buf.data() + pos, buf.size() - pos, k);
        if (k.keys.empty() || k.keys[0].size() != parms.moduli.size() - 1) throw std::logic_error("RelinKeys data is invalid");
        rk.load(ctx, k.flat(0).data());  // index 0 = key power 2 (RelinKeys::get_index)
    }
    Ciphertext load_ciphertext(const SEALContext &ctx)
    {
        he::wire::CtData c;
        pos += he::wire::load_ciphertext(parms, buf.data() + pos, buf.size() - pos, c);
        if (!c.is_ntt_form) throw std::invalid_argument("CKKS encrypted must be in NTT form");
        Ciphertext ct;
        ct.load(ctx, c.data.data(), (std::uint32_t)c.size, (std::uint32_t)c.limbs, c.scale);
        return ct;
    }
    void save_ciphertext(bytes &out, const Ciphertext &ct) const
    {
        he::wire::CtData c;
        c.size = ct.size();
        c.limbs = ct.coeff_modulus_size();
        c.n = parms.n;
        c.scale = ct.scale();
        c.parms_id = parms.parms_id((std::size_t)c.limbs);
        c.data = ct.save();
        const bytes b = he::wire::save_ciphertext(c);
        out.insert(out.end(), b.begin(), b.end());
    }
};
}  // namespace

bytes server_side_simple(const bytes &request, int device)
{
    Session s(request);
    SEALContext ctx = s.context(device);
    RelinKeys rk;
    s.load_relin_keys(ctx, rk);
    Ciphertext op1_ct = s.load_ciphertext(ctx), op2_ct = s.load_ciphertext(ctx);
    Evaluator eval(ctx);
    Ciphertext res_ct = eval % op1_ct * op2_ct;
    res_ct &= eval % rk;  // relin
    res_ct ^= eval;       // rescale
    bytes out;
    s.save_ciphertext(out, res_ct);
    return out;
}

bytes server_side_batch_matmul(const bytes &request, std::size_t r1, std::size_t c1, std::size_t r2, std::size_t c2, int device)
{
    Session s(request);
    SEALContext ctx = s.context(device);
    RelinKeys rk;
    s.load_relin_keys(ctx, rk);
    std::vector<Ciphertext> e1(r1 * c1), e2(r2 * c2);
    for (auto &c : e1) c = s.load_ciphertext(ctx);
    for (auto &c : e2) c = s.load_ciphertext(ctx);
    Evaluator eval(ctx);
    he::linalg::Matrix mat1_ct(r1, c1, std::move(e1)), mat2_ct(r2, c2, std::move(e2));
    he::linalg::Matrix mat3_ct = mat1_ct.matmul(eval, rk, mat2_ct);
    bytes out;
    for (const auto &c : mat3_ct.get_elems()) s.save_ciphertext(out, c);
    return out;
}

bytes server_side_fft(const bytes &request, std::size_t vec_elem_no, int device)
{
    Session s(request);
    SEALContext ctx = s.context(device);
    std::vector<Ciphertext> vec_ct(vec_elem_no);
    for (auto &c : vec_ct) c = s.load_ciphertext(ctx);
    Evaluator eval(ctx);
    CKKSEncoder cencd(ctx);
    std::vector<Ciphertext> vecr_ct = he::fft::fft(cencd, eval, vec_ct);
    bytes out;
    for (const auto &c : vecr_ct) s.save_ciphertext(out, c);
    return out;
}

}  // namespace he::server
