/*
 * hegpu.h -- C ABI of the B200-native CKKS evaluator (libhegpu.so).
 *
 * This is the drop-in boundary for the encrypted linear-algebra hot path of
 * isteiakakis/Homomorphic-Encryption-Algorithms-Diploma-Thesis.  The reference has no
 * FFI: its boundary is the non-virtual class `seal::Evaluator`, passed by const& into
 * every routine (reference include/he_linalg.h:144,406, include/he_fft.h:12-27,
 * include/he_operators.h:44-159).  Each entry point below names the seal::Evaluator
 * method (and the reference call site) it replaces; INTEGRATION.md shows the adapter a
 * maintainer adds on the SEAL side.
 *
 * Conventions
 *  - plain C, no exceptions: every function returns a hegpu_status; hegpu_last_error()
 *    returns the message (SEAL's wording where SEAL would have thrown) for the calling
 *    thread.  HEGPU_ERR_INVALID_ARGUMENT <-> std::invalid_argument,
 *    HEGPU_ERR_LOGIC <-> std::logic_error.
 *  - host buffers use SEAL's layouts: ciphertext uint64_t[size][L][N], plaintext
 *    uint64_t[L][N] (NTT form), key-switching key uint64_t[Lmax][2][K][N]; all values are
 *    canonical residues.  L = number of RNS limbs at the ciphertext's level, K = Lmax+1
 *    key-level limbs (last = special prime).
 *  - a hegpu_ct is a *batch* of B ciphertexts with identical metadata, resident in HBM as
 *    uint64_t[B][size_cap][L_cap][N].  All primitives act on the whole batch.
 *  - one CUDA stream per context; calls are asynchronous on that stream unless they copy
 *    to pageable host memory.  One host thread per context.
 *  - there is no CPU fallback anywhere behind this interface.
 */
#ifndef HEGPU_H
#define HEGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    HEGPU_OK = 0,
    HEGPU_ERR_INVALID_ARGUMENT = 1, /* SEAL: std::invalid_argument */
    HEGPU_ERR_LOGIC = 2,            /* SEAL: std::logic_error      */
    HEGPU_ERR_CUDA = 3,
    HEGPU_ERR_OUT_OF_MEMORY = 4
} hegpu_status;

typedef struct hegpu_ctx hegpu_ctx;
typedef struct hegpu_ct hegpu_ct; /* batch of ciphertexts in HBM */
typedef struct hegpu_pt hegpu_pt; /* set of plaintexts in HBM     */

const char *hegpu_last_error(void);
const char *hegpu_version(void);

/* ---- context: replaces seal::SEALContext + the NTT / RNS tables under seal::Evaluator.
 * moduli[0..K-1] is the coeff_modulus chain (CoeffModulus::Create order), last = special
 * prime.  Roots are SEAL's minimal primitive 2N-th roots.  n in {4096, 8192, 16384, 32768}. */
int hegpu_ctx_create(hegpu_ctx **ctx, uint32_t n, const uint64_t *moduli, uint32_t K, int device);
int hegpu_ctx_destroy(hegpu_ctx *ctx);
int hegpu_sync(hegpu_ctx *ctx);
void *hegpu_ctx_stream(hegpu_ctx *ctx); /* cudaStream_t */
uint64_t hegpu_ctx_psi(hegpu_ctx *ctx, uint32_t mod_index);
uint64_t hegpu_launch_count(hegpu_ctx *ctx); /* kernels launched so far on this context */

/* ---- keys: seal::RelinKeys (index 0) / seal::GaloisKeys::data()[(elt-1)/2], SEAL layout. */
int hegpu_load_relin_key(hegpu_ctx *ctx, const uint64_t *host_key);
int hegpu_load_galois_key(hegpu_ctx *ctx, uint32_t galois_elt, const uint64_t *host_key);
int hegpu_has_galois_key(hegpu_ctx *ctx, uint32_t galois_elt);
/* GaloisTool::get_elt_from_step */
int hegpu_galois_elt_from_step(hegpu_ctx *ctx, int step, uint32_t *elt);

/* ---- ciphertext batches (seal::Ciphertext x B) */
int hegpu_ct_create(hegpu_ctx *ctx, hegpu_ct **ct, uint32_t batch, uint32_t size_cap, uint32_t L_cap);
int hegpu_ct_destroy(hegpu_ct *ct);
/* host [batch][size][L][N]; sets the batch's metadata */
int hegpu_ct_upload(hegpu_ct *ct, const uint64_t *host, uint32_t size, uint32_t L, double scale);
int hegpu_ct_download(hegpu_ct *ct, uint64_t *host);
int hegpu_ct_upload_one(hegpu_ct *ct, uint32_t index, const uint64_t *host);
int hegpu_ct_download_one(hegpu_ct *ct, uint32_t index, uint64_t *host);
/* Asynchronous variants for pipelining (pinned host memory, packed layout size == size_cap and
 * L == L_cap): the copy runs on a dedicated copy stream, starts after all compute enqueued
 * before the call and overlaps compute enqueued after it; later compute that touches the batch
 * waits for the copy on the device.  hegpu_ct_copy_wait blocks the host until the batch's last
 * asynchronous copy has finished (call it before reading a downloaded host buffer). */
int hegpu_ct_upload_async(hegpu_ct *ct, const uint64_t *host, uint32_t size, uint32_t L, double scale);
int hegpu_ct_download_async(hegpu_ct *ct, uint64_t *host);
int hegpu_ct_copy_wait(hegpu_ct *ct);
int hegpu_ct_info(const hegpu_ct *ct, uint32_t *batch, uint32_t *size, uint32_t *L, double *scale);
int hegpu_ct_set_scale(hegpu_ct *ct, double scale);
/* declare what a batch holds after its device buffer was written through hegpu_ct_device_view (e.g. by an NCCL
 * receive): `size` polynomials at level L with the given scale, strides as reported by hegpu_ct_device_view */
int hegpu_ct_set_meta(hegpu_ct *ct, uint32_t size, uint32_t L, double scale);
int hegpu_ct_copy(hegpu_ctx *ctx, hegpu_ct *dst, const hegpu_ct *src);
/* copy src[src_index] into dst[dst_index] (metadata must agree or dst is uninitialised) */
int hegpu_ct_copy_one(hegpu_ctx *ctx, hegpu_ct *dst, uint32_t dst_index, const hegpu_ct *src, uint32_t src_index);
/* raw device view for NCCL / torch interop: base pointer and strides in uint64 elements */
int hegpu_ct_device_view(hegpu_ct *ct, void **dptr, size_t *batch_stride, size_t *poly_stride, size_t *limb_stride);

/* ---- plaintext sets (seal::Plaintext x count, NTT form at level L) */
int hegpu_pt_create(hegpu_ctx *ctx, hegpu_pt **pt, uint32_t count, uint32_t L_cap);
int hegpu_pt_destroy(hegpu_pt *pt);
int hegpu_pt_upload(hegpu_pt *pt, const uint64_t *host, uint32_t L, double scale); /* [count][L][N] */
/* extended plaintexts for HEGPU_MATVEC_DH: host [count][L+1][N], limbs 0..L-1 as above and limb L =
 * the same integer polynomial mod the special prime (NTT form) -- CKKSEncoder::encode at the key
 * level restricted to the limbs q_0..q_{L-1}, P.  Needs L_cap >= L+1. */
int hegpu_pt_upload_ext(hegpu_pt *pt, const uint64_t *host, uint32_t L, double scale);
int hegpu_pt_upload_one(hegpu_pt *pt, uint32_t index, const uint64_t *host);
int hegpu_pt_download_one(hegpu_pt *pt, uint32_t index, uint64_t *host);

/* ---- evaluator primitives.  `out` may alias an input.  `b` (second operand) may have
 * batch 1, which broadcasts it over a's batch.  pt_index >= 0 selects one plaintext of
 * the set for the whole batch; pt_index = -1 pairs plaintext i with ciphertext i.
 *
 *   hegpu_negate             Evaluator::negate[_inplace]            he_operators.cpp:16,26
 *   hegpu_add / hegpu_sub    Evaluator::add / sub[_inplace]         he_operators.cpp:35,45,73,83
 *   hegpu_add_plain/sub_plain Evaluator::add_plain / sub_plain      he_operators.cpp:54,64,92,102
 *   hegpu_multiply           Evaluator::multiply[_inplace]          he_operators.cpp:111,121
 *   hegpu_square             Evaluator::square[_inplace]            he_linalg.cpp:647
 *   hegpu_multiply_plain     Evaluator::multiply_plain[_inplace]    he_operators.cpp:130,140; he_util.h:35,43
 *   hegpu_relinearize        Evaluator::relinearize[_inplace]       he_operators.cpp:149,159
 *   hegpu_rescale_to_next    Evaluator::rescale_to_next[_inplace]   he_operators.cpp:168,178; he_util.h:36,44
 *   hegpu_mod_switch_to_next Evaluator::mod_switch_to_next[_inplace] he_operators.cpp:187,197
 *   hegpu_rotate_vector      Evaluator::rotate_vector[_inplace]     he_operators.cpp:206-235; he_linalg.cpp:595-636
 *   hegpu_apply_galois       Evaluator::apply_galois[_inplace]      (under rotate_vector)
 */
int hegpu_negate(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a);
int hegpu_add(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a, const hegpu_ct *b);
int hegpu_sub(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a, const hegpu_ct *b);
int hegpu_add_plain(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a, const hegpu_pt *pt, int pt_index);
int hegpu_sub_plain(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a, const hegpu_pt *pt, int pt_index);
int hegpu_multiply_plain(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a, const hegpu_pt *pt, int pt_index);
int hegpu_multiply(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a, const hegpu_ct *b);
int hegpu_square(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a);
int hegpu_relinearize(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a);
int hegpu_rescale_to_next(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a);
int hegpu_mod_switch_to_next(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a);
int hegpu_rotate_vector(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a, int steps);
int hegpu_apply_galois(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a, uint32_t galois_elt);

/* ---- kernels exposed for measurement: `count` limb polynomials
 * [count][N] in device memory; polynomial i uses modulus (first_mod + i % n_mods). */
int hegpu_ntt_forward_device(hegpu_ctx *ctx, void *d_data, uint32_t count, uint32_t first_mod, uint32_t n_mods);
int hegpu_ntt_inverse_device(hegpu_ctx *ctx, void *d_data, uint32_t count, uint32_t first_mod, uint32_t n_mods);
/* same on host buffers (upload, transform, download) -- used by host-side tooling */
int hegpu_ntt_forward_host(hegpu_ctx *ctx, uint64_t *host, uint32_t count, uint32_t first_mod, uint32_t n_mods);
int hegpu_ntt_inverse_host(hegpu_ctx *ctx, uint64_t *host, uint32_t count, uint32_t first_mod, uint32_t n_mods);

/* Issue-rate microbenchmarks on this device (bench.py's integer / FP64 rooflines): kind 0 = IMAD.WIDE.U32 (32x32->64
 * multiply-add), 1 = DFMA, 2 = IMAD (32-bit); and of the product's own arithmetic sequences on register operands:
 * 3 = one 64x64->128-bit multiply-accumulate (the unit of the key inner products), 4 = one 60-bit forward NTT
 * butterfly (Shoup), 5 = one FP64 NTT butterfly (modulus 1 of the chain must be below 2^43).
 * *ops_per_second = thread-level operations per second with every SM busy.  Blocking. */
int hegpu_pipe_peak(hegpu_ctx *ctx, int kind, double *ops_per_second);

/* ---- composites that stay on the device.
 *
 * hegpu_matvec_bsgs: plaintext-diagonal x encrypted-vector product, baby-step/giant-step,
 * replacing the per-diagonal loop he_linalg.cpp:977-1003 (case A, p = 1) and the stage
 * body he_fft.cpp:178-203 for plaintext diagonals:
 *   out_b = rescale( sum_g rot_{g*n1}( sum_k pt[g*n1+k] (.) rot_k(in_b) ) ),  k < n1, g < n2
 * diags holds n1*n2 plaintexts, diagonal g*n1+k pre-rotated right by g*n1 slots.
 * Needs Galois keys for steps 1..n1-1 and g*n1 (g = 1..n2-1).
 * flags: HEGPU_MATVEC_RESCALE  apply the final rescale (clear it when multi-GPU partial sums are
 *                              reduced first);
 *        HEGPU_MATVEC_HOIST    hoisted baby steps: the baby-step rotations share one
 *                              digit decomposition (INTT + lift once per ciphertext, the Galois
 *                              permutation is applied to the lifted digits);
 *        HEGPU_MATVEC_LAZY     the giant-step key-switches are summed in the extended basis and
 *                              share ONE mod-down.
        HEGPU_MATVEC_DH       double-hoisted (Bossuat et al., EUROCRYPT 2021): one digit decomposition
 *                              for all baby steps, the rotated ciphertexts stay in the extended basis
 *                              q_0..q_{L-1},P scaled by P (no mod-down per baby step), the inner sums
 *                              are taken there against plaintexts that carry a limb mod P
 *                              (hegpu_pt_upload_ext); of a rotated giant step only component 1 (the one that
 *                              is key-switched) is divided by P, component 0 is permuted in the extended
 *                              basis (the Galois map acts limb-wise, also mod P) and added to the sum of the
 *                              key inner products, and everything shares ONE final mod-down.
 *                              n1 <= 32.  Implies HOIST and LAZY.
 * HOIST / LAZY / DH compute the same function up to key-switch noise but different bits than a
 * chain of rotate_vector calls; with neither flag the composite is exactly the chain of SEAL
 * primitives.  hegpu_matvec_bsgs_range is the diagonal-sharded form: this call
 * owns the n2 giant steps g_first .. g_first+n2-1 (diags holds their n1*n2 diagonals); the
 * partial results of all ranks are summed (NCCL uint64 sum + hegpu_reduce_fixup) before the
 * rescale.  Without LAZY the sum over ranks is bit-identical to the single-call result.
 */
#define HEGPU_MATVEC_RESCALE 1
#define HEGPU_MATVEC_HOIST 2
#define HEGPU_MATVEC_LAZY 4
#define HEGPU_MATVEC_DH 8
/* EXPERIMENTAL, opt-in (with HEGPU_MATVEC_DH, 3-limb level): the fused inner sums run as exact 8-bit-limb integer
 * matrix products on the warp-level integer MMA units (diagonals pre-multiplied with the baby-step keys once per
 * key set); same bits as without the flag.  Not the default: the north star keeps this path off the tensor cores. */
#define HEGPU_MATVEC_IMMA 16
int hegpu_matvec_bsgs(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *in, const hegpu_pt *diags, uint32_t n1,
                      uint32_t n2, int flags);
int hegpu_matvec_bsgs_range(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *in, const hegpu_pt *diags, uint32_t n1,
                            uint32_t n2, uint32_t g_first, int flags);

/* hegpu_bmatmul_diag: BatchedMatrix::matmul (he_linalg.cpp:943-1006) in the reference's
 * own loop order ("exact" mode): res_i = rescale(relin( sum_j rot(x_{xi(i,j)}, steps(i,j)) * d_j )).
 * diag_cts: batch n (the `this` operand), vec_cts: batch p or n (the `other` operand).
 * case_b = 0: case A (this = diagonals, other = columns; x index i, steps j);
 * case_b = 1: case B (this = columns, other = transposed columns; x index j, steps i).
 * out: batch p, size 2, level L-1.  Rotations follow SEAL's NAF decomposition. */
int hegpu_bmatmul(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *this_cts, const hegpu_ct *other_cts, uint32_t n,
                  uint32_t p, int case_b);

/* hegpu_matmul_elemwise: Matrix::matmul (he_linalg.cpp:202-236): one ciphertext per
 * entry, column-major; a is rows x inner, b is inner x cols; out rows x cols at level L-1:
 * C(i,j) = rescale(relin(sum_k A(i,k) * B(k,j))).  a_transposed / b_transposed mirror the
 * reference's transposed flag (he_linalg.cpp:376-384). */
int hegpu_matmul_elemwise(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a, const hegpu_ct *b, uint32_t rows,
                          uint32_t inner, uint32_t cols, int a_transposed, int b_transposed);

/* hegpu_bfft_stage: one stage of he::fft::bfft_ (he_fft.cpp:178-203):
 *   y <- rescale(y*D0) + rescale(rot(y,+steps)*D1) [+ rescale(rot(y,-steps)*D2)]
 * stage_pts holds D0, D1, D2 (count 3; D2 ignored when with_d2 = 0) encoded by the caller
 * at y's level and scale (diag_D, he_fft.cpp:89-164, stays on the host). */
int hegpu_bfft_stage(hegpu_ctx *ctx, hegpu_ct *y, const hegpu_pt *stage_pts, int steps, int with_d2);

/* hegpu_fft_butterflies: one recursion level of he::fft::fft_ (he_fft.cpp:54-65) over a
 * batch: for k < half: t = rescale(odd_k * w_k); e = rescale(even_k * one);
 * out_k = e + t; out_{k+half} = e - t.  even/odd: batch `half`; w_pts: `half` plaintexts;
 * one_pt: 1 plaintext; out: batch 2*half. */
int hegpu_fft_butterflies(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *even, const hegpu_ct *odd,
                          const hegpu_pt *w_pts, const hegpu_pt *one_pt);

/* ---- per-kernel timing for bench.py's roofline (CUDA events on the context's stream around
 * every launch while enabled).  kind indexes the kernel families; hegpu_profile_kind_name
 * returns the family name or NULL past the end.  units = limb polynomials (NTT kernels) or
 * coefficients (element-wise kernels); algo_bytes = algorithmic HBM bytes. */
int hegpu_profile_enable(hegpu_ctx *ctx, int on);
int hegpu_profile_reset(hegpu_ctx *ctx);
const char *hegpu_profile_kind_name(int kind);
int hegpu_profile_read(hegpu_ctx *ctx, int kind, double *ms_total, uint64_t *launches, uint64_t *units,
                       uint64_t *algo_bytes);

/* ---- Ciphertext::is_transparent (SEAL 4.1 ciphertext.h; checked after every Evaluator operation when SEAL is
 * built with SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT, its default: std::logic_error("result ciphertext is
 * transparent")).  *count = number of ciphertexts of the batch whose polynomials 1..size-1 are all zero.
 * Blocking (one reduction kernel + a device-to-host read): the adapter calls it after an operation only when
 * the check is switched on (host/hegpu_seal_like.hpp: Evaluator::throw_on_transparent). */
int hegpu_ct_transparent(hegpu_ctx *ctx, const hegpu_ct *ct, uint32_t *count);

/* ---- multi-GPU: after an NCCL uint64 sum of `terms` partial ciphertexts the
 * residues are < terms*q; reduce them back to [0,q). */
int hegpu_reduce_fixup(hegpu_ctx *ctx, hegpu_ct *ct, uint32_t terms);
/* The same fix-up folded into the rescale that follows it ("a fused fix-up kernel ... as the first step
 * of rescale"): `a` holds the element-wise uint64 sum of `terms` <= 16 partial ciphertexts at level L, out =
 * rescale_to_next(a mod q) at level L-1.  The reduction rides in the operand loads of the rescale's two transforms,
 * so the summed batch crosses HBM once instead of three times.  out must not be `a`. */
int hegpu_rescale_sum_to_next(hegpu_ctx *ctx, hegpu_ct *out, const hegpu_ct *a, uint32_t terms);

#ifdef __cplusplus
}
#endif
#endif
