// seal_golden.cpp -- dump real Microsoft SEAL 4.1 evaluator vectors for the parity pin (tests/test_seal_pin.py).
//
// NOT part of the product and not built by __graft_entry__.build(): SEAL is absent from this image.  A maintainer
// with SEAL 4.1 installed builds and runs it once and drops the output where the tests look for it:
//
//   g++ -std=c++17 -O2 tools/seal_golden.cpp -I<seal>/include/SEAL-4.1 -L<seal>/lib -lseal-4.1 -o seal_golden
//   ./seal_golden tests/golden/seal_vectors_8192.hegvec  8192  60 40 40 60
//   ./seal_golden tests/golden/seal_vectors_16384.hegvec 16384 60 40 40 60
//
// Every array is the raw RNS limb data SEAL holds (Ciphertext::data(), KSwitchKeys::data()[idx][j].data(),
// SecretKey::data()), so the harness feeds the SAME keys and input ciphertexts to the oracle and to the GPU
// evaluator and compares their outputs with SEAL's bit for bit.  The operations are the ones the reference's
// he::operators call (src/core/he_operators.cpp:14-237): add, sub, negate, multiply, square, relinearize,
// rescale_to_next, mod_switch_to_next, multiply_plain, add_plain, sub_plain, rotate_vector (own key and NAF chain).
//
// File format "HEGVEC1" (little endian): magic[8] = "HEGVEC1\0", u32 records, then per record
//   u32 name_len, name, u32 dtype (0 = u64, 1 = f64, 2 = i64), u32 ndim, u64 dims[ndim], payload.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "seal/seal.h"

using namespace seal;

namespace {

struct Record {
    std::string name;
    uint32_t dtype;
    std::vector<uint64_t> dims;
    std::vector<uint64_t> words;  // payload as 8-byte words (f64 / i64 bit patterns)
};

std::vector<Record> records;

void put_u64(const std::string &name, const std::vector<uint64_t> &dims, const uint64_t *p)
{
    size_t count = 1;
    for (auto d : dims) count *= (size_t)d;
    records.push_back(Record{ name, 0, dims, std::vector<uint64_t>(p, p + count) });
}
void put_f64(const std::string &name, double v)
{
    uint64_t w;
    std::memcpy(&w, &v, 8);
    records.push_back(Record{ name, 1, { 1 }, { w } });
}
void put_i64(const std::string &name, const std::vector<int64_t> &v)
{
    Record r{ name, 2, { (uint64_t)v.size() }, {} };
    for (auto x : v) r.words.push_back((uint64_t)x);
    records.push_back(r);
}

// Ciphertext::data() is [size][coeff_modulus_size][N], contiguous
void put_ct(const std::string &name, const Ciphertext &c)
{
    put_u64(name, { (uint64_t)c.size(), (uint64_t)c.coeff_modulus_size(), (uint64_t)c.poly_modulus_degree() }, c.data());
    put_f64(name + ".scale", c.scale());
}
void put_pt(const std::string &name, const Plaintext &p, size_t limbs, size_t n)
{
    put_u64(name, { (uint64_t)limbs, (uint64_t)n }, p.data());
    put_f64(name + ".scale", p.scale());
}
// one key-switching key: data()[idx][j] is a PublicKey whose Ciphertext is [2][K][N]  ->  [K-1][2][K][N]
void put_kswitch(const std::string &name, const std::vector<PublicKey> &key, size_t K, size_t n)
{
    std::vector<uint64_t> flat;
    for (const auto &pk : key) flat.insert(flat.end(), pk.data().data(), pk.data().data() + 2 * K * n);
    put_u64(name, { (uint64_t)key.size(), 2, (uint64_t)K, (uint64_t)n }, flat.data());
}

void write_file(const char *path)
{
    FILE *f = std::fopen(path, "wb");
    if (!f) {
        std::perror(path);
        std::exit(1);
    }
    const char magic[8] = { 'H', 'E', 'G', 'V', 'E', 'C', '1', 0 };
    std::fwrite(magic, 1, 8, f);
    const uint32_t cnt = (uint32_t)records.size();
    std::fwrite(&cnt, 4, 1, f);
    for (const auto &r : records) {
        const uint32_t nl = (uint32_t)r.name.size(), nd = (uint32_t)r.dims.size();
        std::fwrite(&nl, 4, 1, f);
        std::fwrite(r.name.data(), 1, nl, f);
        std::fwrite(&r.dtype, 4, 1, f);
        std::fwrite(&nd, 4, 1, f);
        std::fwrite(r.dims.data(), 8, nd, f);
        std::fwrite(r.words.data(), 8, r.words.size(), f);
    }
    std::fclose(f);
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s out.hegvec N bits...   (e.g. 8192 60 40 40 60)\n", argv[0]);
        return 2;
    }
    const size_t n = (size_t)std::atoll(argv[2]);
    std::vector<int> bits;
    for (int i = 3; i < argc; ++i) bits.push_back(std::atoi(argv[i]));
    const double scale = std::ldexp(1.0, bits[1]);

    EncryptionParameters parms(scheme_type::ckks);
    parms.set_poly_modulus_degree(n);
    parms.set_coeff_modulus(CoeffModulus::Create(n, bits));
    SEALContext context(parms, true, sec_level_type::none);  // the reference uses tc128 chains; `none` also admits the cfg 4 test chain
    const size_t K = parms.coeff_modulus().size(), L = K - 1;

    KeyGenerator keygen(context);
    const SecretKey &sk = keygen.secret_key();
    PublicKey pk;
    keygen.create_public_key(pk);
    RelinKeys rk;
    keygen.create_relin_keys(rk);
    GaloisKeys gk;
    // a subset of the default power-of-two set (reference demos: create_galois_keys with no list,
    // src/demos/matrix_operations.cpp:1060), enough for the rotations below and small enough to commit the file
    keygen.create_galois_keys(std::vector<int>{ 1, -1, 2, 4, -4, 8, 64, (int)(n / 4) }, gk);
    Encryptor encryptor(context, pk);
    Evaluator ev(context);
    CKKSEncoder encoder(context);

    std::vector<uint64_t> moduli;
    for (const auto &m : parms.coeff_modulus()) moduli.push_back(m.value());
    put_u64("moduli", { (uint64_t)K }, moduli.data());
    put_i64("bits", std::vector<int64_t>(bits.begin(), bits.end()));
    put_i64("n", { (int64_t)n });
    put_u64("secret_key", { (uint64_t)K, (uint64_t)n }, sk.data().data());
    put_kswitch("relin_key", rk.data()[RelinKeys::get_index(2)], K, n);
    {
        std::vector<int64_t> elts;
        for (size_t idx = 0; idx < gk.data().size(); ++idx)
            if (!gk.data()[idx].empty()) {
                const uint32_t elt = (uint32_t)(2 * idx + 1);
                elts.push_back(elt);
                put_kswitch("galois_key." + std::to_string(elt), gk.data()[idx], K, n);
            }
        put_i64("galois_elts", elts);
    }

    // inputs: slot i of a is (i % 17) * 0.25 - 2, of b is 1.5 - (i % 5), plain is (i % 7) * 0.5
    const size_t slots = n / 2;
    std::vector<double> va(slots), vb(slots), vp(slots);
    for (size_t i = 0; i < slots; ++i) {
        va[i] = (double)(i % 17) * 0.25 - 2.0;
        vb[i] = 1.5 - (double)(i % 5);
        vp[i] = (double)(i % 7) * 0.5;
    }
    Plaintext pa, pb, pp;
    encoder.encode(va, scale, pa);
    encoder.encode(vb, scale, pb);
    encoder.encode(vp, scale, pp);
    Ciphertext a, b;
    encryptor.encrypt(pa, a);
    encryptor.encrypt(pb, b);
    put_ct("ct_a", a);
    put_ct("ct_b", b);
    put_pt("pt", pp, L, n);

    Ciphertext r;
    ev.add(a, b, r);
    put_ct("add", r);
    ev.sub(a, b, r);
    put_ct("sub", r);
    ev.negate(a, r);
    put_ct("negate", r);
    ev.add_plain(a, pp, r);
    put_ct("add_plain", r);
    ev.sub_plain(a, pp, r);
    put_ct("sub_plain", r);
    ev.multiply_plain(a, pp, r);
    put_ct("multiply_plain", r);
    ev.mod_switch_to_next(a, r);
    put_ct("mod_switch", r);

    Ciphertext m3, sq3, m2, m2r;
    ev.multiply(a, b, m3);
    put_ct("multiply", m3);
    ev.square(a, sq3);
    put_ct("square", sq3);
    ev.relinearize(m3, rk, m2);
    put_ct("relinearize", m2);
    ev.rescale_to_next(m2, m2r);
    put_ct("rescale", m2r);

    // rotations at the top level and one level down; 7 and -5 have no key of their own in the default set -> NAF chains
    const std::vector<int64_t> steps = { 1, -1, 2, 64, 7, -5, (int64_t)(slots / 2), 0 };
    put_i64("rotate_steps", steps);
    for (auto st : steps) {
        ev.rotate_vector(a, (int)st, gk, r);
        put_ct("rotate." + std::to_string(st), r);
        ev.rotate_vector(m2r, (int)st, gk, r);
        put_ct("rotate_low." + std::to_string(st), r);
    }
    write_file(argv[1]);
    std::printf("%zu records -> %s (N = %zu, K = %zu)\n", records.size(), argv[1], n, K);
    return 0;
}
