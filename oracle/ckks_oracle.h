/*
 * ckks_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the Microsoft SEAL 4.1 CKKS arithmetic that the
 * reference (isteiakakis/Homomorphic-Encryption-Algorithms-Diploma-Thesis)
 * reaches through `seal::Evaluator` on its he_linalg / he_fft / he_operators
 * hot path (reference call sites: src/core/he_operators.cpp:14-237,
 * src/core/he_linalg.cpp:595-647,943-1006, src/core/he_fft.cpp:13-223,
 * include/he_util.h:27-77).
 *
 * PARITY UNPINNED: SEAL 4.1 is an un-vendored external dependency
 * (reference CMakeLists.txt:22-24) and is absent from this machine, and the
 * reference ships no tests / golden vectors for this path.  The algorithms
 * below restate SEAL 4.1's published algorithms (evaluator.cpp
 * switch_key_inplace / rotate_internal / apply_galois_inplace /
 * ckks_multiply / multiply_plain_ntt / mod_switch_scale_to_next,
 * util/rns.cpp divide_and_round_q_last_ntt_inplace, util/galois.cpp,
 * util/ntt.cpp, util/numth.cpp get_primes / try_minimal_primitive_root / naf,
 * modulus.cpp CoeffModulus::Create, keygenerator.cpp
 * generate_one_kswitch_key).  See SURVEY.md section 9 for the spec.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library.  The product (libhegpu.so)
 * never links, loads or calls it.
 *
 * Layout conventions (SEAL's): a ciphertext of `size` polynomials at level L
 * is uint64_t[size][L][N], every value a canonical residue in [0, q_i), NTT
 * form (bit-reversed evaluation order).  A key-switching key is
 * uint64_t[Lmax][2][K][N] (digit, component, key-level limb, coefficient).
 */
#ifndef CKKS_ORACLE_H
#define CKKS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

/* --- parameters ------------------------------------------------------- */
/* SEAL util::get_primes: `count` primes of `bits` bits, = 1 mod factor, descending. */
int orc_get_primes(uint64_t factor, int bits, uint32_t count, uint64_t *out);
/* SEAL CoeffModulus::Create(n, bit_sizes). */
int orc_coeff_modulus_create(uint32_t n, const int *bits, uint32_t count, uint64_t *out);
int orc_is_prime(uint64_t v);

/* moduli[0..K-1], last one is the special prime P. */
orc_ctx *orc_ctx_create(uint32_t n, const uint64_t *moduli, uint32_t K);
void orc_ctx_free(orc_ctx *ctx);
uint32_t orc_n(const orc_ctx *ctx);
uint32_t orc_K(const orc_ctx *ctx);
uint64_t orc_modulus(const orc_ctx *ctx, uint32_t i);
/* minimal primitive 2N-th root of unity mod q_i (SEAL try_minimal_primitive_root) */
uint64_t orc_psi(const orc_ctx *ctx, uint32_t i);

/* --- NTT ------------------------------------------------- */
void orc_ntt_fwd(const orc_ctx *ctx, uint32_t mod_index, uint64_t *a);
void orc_ntt_inv(const orc_ctx *ctx, uint32_t mod_index, uint64_t *a);

/* --- element-wise evaluator ops -------------------------- */
void orc_negate(const orc_ctx *ctx, uint32_t L, const uint64_t *a, uint32_t size, uint64_t *out);
void orc_add(const orc_ctx *ctx, uint32_t L, const uint64_t *a, uint32_t sa, const uint64_t *b, uint32_t sb, uint64_t *out);
void orc_sub(const orc_ctx *ctx, uint32_t L, const uint64_t *a, uint32_t sa, const uint64_t *b, uint32_t sb, uint64_t *out);
void orc_add_plain(const orc_ctx *ctx, uint32_t L, const uint64_t *ct, uint32_t size, const uint64_t *pt, uint64_t *out);
void orc_sub_plain(const orc_ctx *ctx, uint32_t L, const uint64_t *ct, uint32_t size, const uint64_t *pt, uint64_t *out);
void orc_multiply_plain(const orc_ctx *ctx, uint32_t L, const uint64_t *ct, uint32_t size, const uint64_t *pt, uint64_t *out);
void orc_multiply(const orc_ctx *ctx, uint32_t L, const uint64_t *a, uint32_t sa, const uint64_t *b, uint32_t sb, uint64_t *out);
void orc_square(const orc_ctx *ctx, uint32_t L, const uint64_t *a, uint32_t sa, uint64_t *out);

/* --- level management ------------------------------------ */
void orc_rescale(const orc_ctx *ctx, uint32_t L, const uint64_t *ct, uint32_t size, uint64_t *out);
void orc_mod_switch(const orc_ctx *ctx, uint32_t L, const uint64_t *ct, uint32_t size, uint64_t *out);

/* --- Galois / key switching -------------------------- */
uint32_t orc_galois_elt_from_step(uint32_t n, int step);
int orc_naf(int value, int *out);
void orc_galois_table(uint32_t n, uint32_t elt, uint32_t *table);
void orc_apply_galois_ntt(const orc_ctx *ctx, uint32_t limbs, uint32_t elt, const uint64_t *in, uint64_t *out);
void orc_switch_key(const orc_ctx *ctx, uint32_t L, uint64_t *ct, const uint64_t *target, const uint64_t *key);
void orc_relinearize(const orc_ctx *ctx, uint32_t L, const uint64_t *ct3, const uint64_t *rk, uint64_t *out);
void orc_apply_galois(const orc_ctx *ctx, uint32_t L, const uint64_t *ct, uint32_t elt, const uint64_t *key, uint64_t *out);
/* rotate_vector with SEAL's NAF fallback.  keys[i] is the key for elts[i].
 * Returns the number of key-switches performed, or -1 = "Galois key not present". */
int orc_rotate(const orc_ctx *ctx, uint32_t L, const uint64_t *ct, int steps, uint32_t n_keys,
               const uint32_t *elts, const uint64_t *const *keys, uint64_t *out);

/* --- client side (test fixtures: SEAL KeyGenerator / Encryptor / Decryptor) */
void orc_sample_secret(const orc_ctx *ctx, uint64_t seed, uint64_t *s);
void orc_gen_kswitch_key(const orc_ctx *ctx, uint64_t seed, const uint64_t *s, const uint64_t *new_key, uint64_t *out);
void orc_gen_relin_key(const orc_ctx *ctx, uint64_t seed, const uint64_t *s, uint64_t *out);
void orc_gen_galois_key(const orc_ctx *ctx, uint64_t seed, const uint64_t *s, uint32_t elt, uint64_t *out);
void orc_encrypt_symmetric(const orc_ctx *ctx, uint64_t seed, uint32_t L, const uint64_t *s, const uint64_t *plain, uint64_t *ct);
void orc_decrypt(const orc_ctx *ctx, uint32_t L, const uint64_t *ct, uint32_t size, const uint64_t *s, uint64_t *plain);

/* --- composites (restated with the same algorithm the CUDA fast path uses) */
/* BSGS diagonal matvec over a batch of B ciphertexts (OpenMP over the batch):
 *   out_b = rescale( sum_g rot_{g*n1}( sum_b pt[g*n1+b] (.) rot_b(ct_b) ) )
 * pts: [n1*n2][L][N] pre-rotated plaintext diagonals; baby_keys[b] (b=1..n1-1)
 * is the Galois key of step b, giant_keys[g] (g=1..n2-1) of step g*n1.
 * cts: [B][2][L][N];  out: [B][2][L-1][N]. */
void orc_matvec_bsgs(const orc_ctx *ctx, uint32_t L, uint32_t B, const uint64_t *cts, uint32_t n1, uint32_t n2,
                     const uint64_t *pts, const uint64_t *const *baby_keys, const uint64_t *const *giant_keys,
                     uint64_t *out, int threads);
/* restatement of hegpu_matvec_bsgs_range: flags 1 = hoisted baby steps, 2 = lazy giant steps (one
 * mod-down), 4 = final rescale; global giant steps g_first .. g_first+n2-1 */
void orc_matvec_bsgs_ex(const orc_ctx *ctx, uint32_t L, uint32_t B, const uint64_t *cts, uint32_t n1, uint32_t n2,
                        uint32_t g_first, const uint64_t *pts, const uint64_t *const *baby_keys,
                        const uint64_t *const *giant_keys, int flags, uint64_t *out, int threads);
/* = orc_matvec_bsgs_ex with flags 1|2|4 and g_first 0 */
void orc_matvec_bsgs_fast(const orc_ctx *ctx, uint32_t L, uint32_t B, const uint64_t *cts, uint32_t n1, uint32_t n2,
                          const uint64_t *pts, const uint64_t *const *baby_keys, const uint64_t *const *giant_keys,
                          uint64_t *out, int threads);
/* restatement of hegpu_matvec_bsgs_range with HEGPU_MATVEC_DH (double-hoisted BSGS: baby rotations
 * stay in the extended basis, one mod-down per giant step).  ptsx: [n1*n2][L+1][N], limb L = residues
 * mod the special prime.  flags: 4 = final rescale. */
void orc_matvec_bsgs_dh(const orc_ctx *ctx, uint32_t L, uint32_t B, const uint64_t *cts, uint32_t n1, uint32_t n2,
                        uint32_t g_first, const uint64_t *ptsx, const uint64_t *const *baby_keys,
                        const uint64_t *const *giant_keys, int flags, uint64_t *out, int threads);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
