// he_host_test.cpp -- driver for tests/test_host_cpp.py: loads keys / ciphertexts / plaintexts
// from a binary case file, runs the host mirror of the reference's routines
// (he_operators / he_linalg / he_fft / he_util) on the GPU evaluator, and dumps every result
// (plus every plaintext the routines encoded) so that the test can replay the same calls on
// the CPU oracle and compare bit for bit.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include <iterator>

#include "he_fft.h"
#include "he_linalg.h"
#include "he_server.hpp"
#include "he_math.h"
#include "he_util.h"

using namespace he::gpu;
using namespace he::operators;
using he::linalg::BatchedMatrix;
using he::linalg::BatchedVector;
using he::linalg::Matrix;

namespace {
template <class T>
T rd(std::ifstream &f)
{
    T v;
    f.read(reinterpret_cast<char *>(&v), sizeof(T));
    return v;
}
std::vector<std::uint64_t> rdv(std::ifstream &f, std::size_t words)
{
    std::vector<std::uint64_t> v(words);
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(words * 8));
    return v;
}
template <class T>
void wr(std::ofstream &f, T v)
{
    f.write(reinterpret_cast<const char *>(&v), sizeof(T));
}
void wr_ct(std::ofstream &f, const Ciphertext &c)
{
    wr<std::uint32_t>(f, (std::uint32_t)c.size());
    wr<std::uint32_t>(f, (std::uint32_t)c.coeff_modulus_size());
    wr<double>(f, c.scale());
    const auto d = c.save();
    f.write(reinterpret_cast<const char *>(d.data()), (std::streamsize)(d.size() * 8));
}
void wr_pt(std::ofstream &f, const Plaintext &p)
{
    wr<std::uint32_t>(f, p.parms_id().limbs);
    wr<double>(f, p.scale());
    const auto d = p.save();
    f.write(reinterpret_cast<const char *>(d.data()), (std::streamsize)(d.size() * 8));
}
}  // namespace

int main(int argc, char **argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: he_host_test <case.bin> <cmd> <out.bin> [args...]\n");
        return 2;
    }
    try {
        std::ifstream f(argv[1], std::ios::binary);
        const std::string cmd = argv[2];
        auto arg = [&](int i) { return i < argc ? std::atoi(argv[i]) : 0; };
        if (cmd.rfind("serve_", 0) == 0) {  // the reference server's modes: argv[1] is the request payload, argv[3] receives the reply
            const he::wire::bytes req((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            he::wire::bytes rep;
            if (cmd == "serve_simple") rep = he::server::server_side_simple(req);
            else if (cmd == "serve_batch_matmul") rep = he::server::server_side_batch_matmul(req, (std::size_t)arg(4), (std::size_t)arg(5), (std::size_t)arg(6), (std::size_t)arg(7));
            else if (cmd == "serve_fft") rep = he::server::server_side_fft(req, (std::size_t)arg(4));
            else throw std::invalid_argument("unknown server mode");
            std::ofstream o(argv[3], std::ios::binary);
            o.write(reinterpret_cast<const char *>(rep.data()), (std::streamsize)rep.size());
            std::printf("ok %zu reply bytes\n", rep.size());
            return 0;
        }
        auto argd = [&](int i) { return i < argc ? std::atof(argv[i]) : 0.0; };
        const std::uint32_t n = rd<std::uint32_t>(f), K = rd<std::uint32_t>(f);
        const auto moduli = rdv(f, K);
        SEALContext ctx(n, moduli);
        Evaluator eval(ctx);
        CKKSEncoder cencd(ctx);
        RelinKeys rk;
        GaloisKeys gk;
        const std::size_t key_words = (std::size_t)(K - 1) * 2 * K * n;
        if (rd<std::uint32_t>(f)) rk.load(ctx, rdv(f, key_words).data());
        for (std::uint32_t i = 0, c = rd<std::uint32_t>(f); i < c; ++i) {
            const std::uint32_t elt = rd<std::uint32_t>(f);
            gk.load(ctx, elt, rdv(f, key_words).data());
        }
        std::vector<Ciphertext> cts(rd<std::uint32_t>(f));
        for (auto &c : cts) {
            const std::uint32_t size = rd<std::uint32_t>(f), L = rd<std::uint32_t>(f);
            const double scale = rd<double>(f);
            c.load(ctx, rdv(f, (std::size_t)size * L * n).data(), size, L, scale);
        }
        std::vector<Plaintext> pts(rd<std::uint32_t>(f));
        for (auto &p : pts) {
            const std::uint32_t L = rd<std::uint32_t>(f);
            const double scale = rd<double>(f);
            p.load(ctx, rdv(f, (std::size_t)L * n).data(), L, scale);
        }

        std::vector<Ciphertext> out;
        std::vector<Plaintext> logged;
        he::fft::encoded_log = &logged;

        if (cmd == "operators") {
            const Ciphertext &a = cts[0], &b = cts[1];
            const Plaintext &p = pts[0];
            Ciphertext t;
            t = a; t -= eval; out.push_back(t);
            out.push_back(-(eval % a));
            t = a; t += eval % b; out.push_back(t);
            out.push_back(eval % a + b);
            t = a; t += eval % p; out.push_back(t);
            out.push_back(eval % a + p);
            t = a; t -= eval % b; out.push_back(t);
            out.push_back(eval % a - b);
            t = a; t -= eval % p; out.push_back(t);
            out.push_back(eval % a - p);
            t = a; t *= eval % b; out.push_back(t);
            Ciphertext prod = eval % a * b;
            out.push_back(prod);
            t = a; t *= eval % p; out.push_back(t);
            out.push_back(eval % a * p);
            t = prod; t &= eval % rk; out.push_back(t);
            Ciphertext rel = eval % rk & prod;
            out.push_back(rel);
            t = rel; t ^= eval; out.push_back(t);
            out.push_back(eval ^ rel);
            t = a; t |= eval; out.push_back(t);
            out.push_back(eval | a);
            t = a; t <<= eval % gk % 3; out.push_back(t);
            out.push_back(eval % gk % a << 1);
            t = a; t >>= eval % gk % 2; out.push_back(t);
            out.push_back(eval % gk % a >> 5);
        } else if (cmd == "bmatmul") {  // args: case_b n p dim
            const bool case_b = arg(4) != 0;
            const std::size_t nn = (std::size_t)arg(5), p = (std::size_t)arg(6), dim = (std::size_t)arg(7);
            std::vector<BatchedVector> tv, ov;
            for (std::size_t i = 0; i < nn; ++i) tv.emplace_back(dim, cts[i]);
            for (std::size_t i = nn; i < cts.size(); ++i) ov.emplace_back(case_b ? p : dim, cts[i]);
            BatchedMatrix T(case_b ? BatchedMatrix::BatchingType::col : BatchedMatrix::BatchingType::diag, std::move(tv));
            BatchedMatrix O(BatchedMatrix::BatchingType::col, std::move(ov));
            if (case_b) O.transp();
            BatchedMatrix R = T.matmul(eval, rk, gk, O);
            for (auto &v : R.get_bvecs()) out.push_back(v.get_bvec());
        } else if (cmd == "matmul") {  // args: rows inner cols a_t b_t   (physical element order)
            const std::size_t rows = (std::size_t)arg(4), inner = (std::size_t)arg(5), cols = (std::size_t)arg(6);
            const bool at = arg(7) != 0, bt = arg(8) != 0;
            std::vector<Ciphertext> ea(cts.begin(), cts.begin() + (std::ptrdiff_t)(rows * inner)), eb(cts.begin() + (std::ptrdiff_t)(rows * inner), cts.end());
            Matrix A(at ? inner : rows, at ? rows : inner, std::move(ea)), B(bt ? cols : inner, bt ? inner : cols, std::move(eb));
            if (at) A.transp();
            if (bt) B.transp();
            Matrix C = A.matmul(eval, rk, B);
            for (auto &e : C.get_elems()) out.push_back(e);
            if (rows == inner && inner == cols && !at) {
                Matrix S = A.matmul_square(eval, rk);
                for (auto &e : S.get_elems()) out.push_back(e);
                Matrix G = A.left_matmul_with_transp(eval, rk);
                for (auto &e : G.get_elems()) out.push_back(e);
                Matrix E = A;
                E *= eval % rk % B;  // element-wise
                for (auto &e : E.get_elems()) out.push_back(e);
            }
        } else if (cmd == "matpow") {  // args: d powr   (Matrix::matmul_pow, he_linalg.cpp:316-349)
            const std::size_t d = (std::size_t)arg(4);
            std::vector<Ciphertext> ea(cts.begin(), cts.begin() + (std::ptrdiff_t)(d * d));
            Matrix A(d, d, std::move(ea));
            try {
                Matrix R = A.matmul_pow(eval, rk, arg(5));
                for (auto &e : R.get_elems()) out.push_back(e);
            } catch (const std::invalid_argument &e) {  // SEAL's own error when the power mixes levels (e.g. 3 = A * A^2)
                std::printf("matpow_error=%s\n", e.what());
            }
        } else if (cmd == "sum_elems") {  // args: dim
            BatchedVector v((std::size_t)arg(4), cts[0]);
            v.sum_elems_inplace(eval, gk);
            out.push_back(v.get_bvec());
            BatchedVector s = BatchedVector((std::size_t)arg(4), cts[0]).square(eval, rk);
            out.push_back(s.get_bvec());
        } else if (cmd == "bfft") {  // args: n inverse
            out.push_back(arg(5) ? he::fft::ibfft(cencd, eval, gk, cts[0], (std::size_t)arg(4)) : he::fft::bfft(cencd, eval, gk, cts[0], (std::size_t)arg(4)));
        } else if (cmd == "fft") {  // args: inverse
            out = arg(4) ? he::fft::ifft(cencd, eval, cts) : he::fft::fft(cencd, eval, cts);
        } else if (cmd == "drop_levels") {  // args: levels
            Ciphertext t = cts[0];
            Plaintext one;
            he::util::drop_chain_levels(ctx, cencd, eval, one, t, (std::size_t)arg(4));
            out.push_back(t);
            logged.push_back(one);
            Ciphertext u = cts[0];
            he::util::reach_chain_level(ctx, cencd, eval, u, t);
            out.push_back(u);
            wr<std::uint32_t>(*new std::ofstream("/dev/null"), (std::uint32_t)he::util::get_chain_index(ctx, t));
        } else if (cmd == "math") {  // args: which(0 signed_inv, 1 inv_sqrt_twice, 2 sqrt, 3 abs) a iter_num
            const int which = arg(4);
            const double a = argd(5);
            const std::size_t iters = (std::size_t)arg(6);
            if (which == 0) out.push_back(he::math::signed_inv(cencd, eval, rk, cts[0], a, iters));
            else if (which == 1) out.push_back(he::math::inv_sqrt_twice(cencd, eval, rk, cts[0], a, iters));
            else if (which == 2) out.push_back(he::math::sqrt(ctx, cencd, eval, rk, cts[0], a, iters));
            else out.push_back(he::math::abs(ctx, cencd, eval, rk, cts[0], a, iters));
        } else if (cmd == "least_squares") {  // args: n   (bench_he_least_squares_2d, src/demos/matrix_operations.cpp:833-1040, statement by statement)
            const std::size_t nn = (std::size_t)arg(4);
            BatchedVector x_ctv(nn, cts[0]), y_ctv(nn, cts[1]);
            Ciphertext sum_x_ct = x_ctv.sum_elems(eval, gk).get_bvec();
            Ciphertext sum_y_ct = y_ctv.sum_elems(eval, gk).get_bvec();
            Ciphertext sum_xx_ct = x_ctv.square(eval, rk).sum_elems(eval, gk).get_bvec();
            Ciphertext sum_xy_ct = (eval % rk % x_ctv * y_ctv).sum_elems(eval, gk).get_bvec();
            // denominator n * sum(xx) - sum(x)^2
            Plaintext n_pt;
            cencd.encode((double)nn, sum_xx_ct.parms_id(), sum_xx_ct.scale(), n_pt);
            Ciphertext n_sum_xx_ct = eval % sum_xx_ct * n_pt;
            n_sum_xx_ct ^= eval;
            Ciphertext sum_x_sqr_ct;
            eval.square(sum_x_ct, sum_x_sqr_ct);
            sum_x_sqr_ct &= eval % rk;
            sum_x_sqr_ct ^= eval;
            Plaintext one_pt;
            cencd.encode(1.0, sum_x_sqr_ct.parms_id(), sum_x_sqr_ct.scale(), one_pt);
            sum_x_sqr_ct *= eval % one_pt;
            sum_x_sqr_ct ^= eval;
            Ciphertext denom_ct = eval % n_sum_xx_ct - sum_x_sqr_ct;
            Plaintext one_pt_;  // slot 0 only (the reference's FIXME: the inverse must not see the other slots)
            cencd.encode(std::vector<double>{ 1 }, denom_ct.parms_id(), denom_ct.scale(), one_pt_);
            denom_ct *= eval % one_pt_;
            denom_ct ^= eval;
            Ciphertext denom_inv_ct = he::math::signed_inv(cencd, eval, rk, denom_ct, 0.05, 6);
            // numerator of a: n * sum(xy) - sum(x) sum(y)
            Ciphertext n_sum_xy_ct = eval % sum_xy_ct * n_pt;
            n_sum_xy_ct ^= eval;
            Ciphertext sum_x_sum_y_ct = eval % sum_x_ct * sum_y_ct;
            sum_x_sum_y_ct &= eval % rk;
            sum_x_sum_y_ct ^= eval;
            sum_x_sum_y_ct *= eval % one_pt;
            sum_x_sum_y_ct ^= eval;
            Ciphertext a_num_ct = eval % n_sum_xy_ct - sum_x_sum_y_ct;
            // numerator of b: sum(y) sum(xx) - sum(x) sum(xy)
            cencd.encode(1.0, sum_y_ct.parms_id(), sum_y_ct.scale(), one_pt);
            Ciphertext sum_y_sum_xx_ct = sum_y_ct;
            sum_y_sum_xx_ct *= eval % one_pt;
            sum_y_sum_xx_ct ^= eval;
            sum_y_sum_xx_ct *= eval % sum_xx_ct;
            sum_y_sum_xx_ct &= eval % rk;
            sum_y_sum_xx_ct ^= eval;
            Ciphertext sum_x_sum_xy_ct = sum_x_ct;
            sum_x_sum_xy_ct *= eval % one_pt;
            sum_x_sum_xy_ct ^= eval;
            sum_x_sum_xy_ct *= eval % sum_xy_ct;
            sum_x_sum_xy_ct &= eval % rk;
            sum_x_sum_xy_ct ^= eval;
            Ciphertext b_num_ct = eval % sum_y_sum_xx_ct - sum_x_sum_xy_ct;
            // a, b
            he::util::reach_chain_level(ctx, cencd, eval, one_pt, std::vector<Ciphertext *>{ &a_num_ct, &b_num_ct }, denom_inv_ct);
            Ciphertext a_ct = eval % a_num_ct * denom_inv_ct;
            a_ct &= eval % rk;
            a_ct ^= eval;
            Ciphertext b_ct = eval % b_num_ct * denom_inv_ct;
            b_ct &= eval % rk;
            b_ct ^= eval;
            for (const Ciphertext *c : { &denom_ct, &denom_inv_ct, &a_num_ct, &b_num_ct, &a_ct, &b_ct }) out.push_back(*c);
        } else if (cmd == "errors") {
            // exception types and messages must be SEAL's
            int ok = 0;
            try { Ciphertext t = cts[0]; t.set_scale(t.scale() * 2); t += eval % cts[1]; } catch (const std::invalid_argument &e) { ok += std::string(e.what()) == "scale mismatch"; }
            try { Ciphertext t = cts[0]; t <<= eval % gk % 1; } catch (const std::invalid_argument &e) { ok += std::string(e.what()) == "Galois key not present"; }
            try { Ciphertext t = cts[0]; for (int i = 0; i < 10; ++i) t ^= eval; } catch (const std::invalid_argument &e) { ok += std::string(e.what()) == "end of modulus switching chain reached"; }
            // SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT: a - a has no polynomial beyond c0 that is non-zero
            {
                Evaluator strict(ctx);
                strict.throw_on_transparent = true;
                try { Ciphertext t = cts[0]; t -= strict % cts[0]; } catch (const std::logic_error &e) { ok += std::string(e.what()) == "result ciphertext is transparent"; }
                Ciphertext t = cts[0];
                t += strict % cts[1];  // an ordinary result passes the check
                ok += 1;
            }
            std::printf("errors_ok=%d\n", ok);
        } else {
            std::fprintf(stderr, "unknown command %s\n", cmd.c_str());
            return 2;
        }
        ctx.sync();
        std::ofstream o(argv[3], std::ios::binary);
        wr<std::uint32_t>(o, (std::uint32_t)out.size());
        for (auto &c : out) wr_ct(o, c);
        wr<std::uint32_t>(o, (std::uint32_t)logged.size());
        for (auto &p : logged) wr_pt(o, p);
        std::printf("ok %zu ciphertexts %zu plaintexts\n", out.size(), logged.size());
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
}
