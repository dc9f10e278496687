// he_linalg.cpp -- host mirror of src/core/he_linalg.cpp over the GPU evaluator.
//
// Element-wise operators: one he::operators call per ciphertext, like the reference.
// Hot-path routines: operands are gathered into device batches and a single composite of the
// C ABI runs the whole product on the device in the reference's loop order
//   Matrix::matmul / left_matmul_with_transp / matmul_square  -> hegpu_matmul_elemwise
//   BatchedMatrix::matmul (case A and case B)                  -> hegpu_bmatmul
// so results are bit-identical to the reference's per-operator loops while the ciphertexts
// never leave HBM between the n*p rotations, products, additions, relinearisations, rescales.
#include "he_linalg.h"

#include <cassert>
#include <cmath>

namespace he::linalg {

using namespace he::operators;
using he::gpu::check;
using he::gpu::SEALContext;

namespace {

// RAII device batch
struct Batch {
    hegpu_ct *h = nullptr;
    Batch() = default;
    Batch(const Batch &) = delete;
    ~Batch() { hegpu_ct_destroy(h); }
};

template <class GetCt>
void gather(const SEALContext &ctx, Batch &b, std::size_t count, std::uint32_t size_cap, GetCt &&get)
{
    check(hegpu_ct_create(ctx.raw(), &b.h, (std::uint32_t)count, size_cap, ctx.key_limbs() - 1));
    for (std::size_t i = 0; i < count; ++i) check(hegpu_ct_copy_one(ctx.raw(), b.h, (std::uint32_t)i, get(i).handle(), 0));
}
template <class PutCt>
void scatter(const SEALContext &ctx, const Batch &b, std::size_t count, PutCt &&put)
{
    for (std::size_t i = 0; i < count; ++i) check(hegpu_ct_copy_one(ctx.raw(), put(i).prepare(ctx), 0, b.h, (std::uint32_t)i));
}

// C = A * B for Matrix operands given as (elements, logical dims, transposed flag)
Matrix device_matmul(const Evaluator &eval, const std::vector<Ciphertext> &a, bool a_t, const std::vector<Ciphertext> &b, bool b_t,
                     std::size_t rows, std::size_t inner, std::size_t cols)
{
    const SEALContext &ctx = eval.context();
    Batch ba, bb, bo;
    gather(ctx, ba, a.size(), 2, [&](std::size_t i) -> const Ciphertext & { return a[i]; });
    gather(ctx, bb, b.size(), 2, [&](std::size_t i) -> const Ciphertext & { return b[i]; });
    check(hegpu_ct_create(ctx.raw(), &bo.h, (std::uint32_t)(rows * cols), 2, ctx.key_limbs() - 1));
    check(hegpu_matmul_elemwise(ctx.raw(), bo.h, ba.h, bb.h, (std::uint32_t)rows, (std::uint32_t)inner, (std::uint32_t)cols, a_t, b_t));
    std::vector<Ciphertext> out(rows * cols);
    scatter(ctx, bo, out.size(), [&](std::size_t i) -> Ciphertext & { return out[i]; });
    return Matrix(rows, cols, std::move(out));
}

}  // namespace

// =============================================================================== Matrix
Matrix::Matrix(std::size_t rows, std::size_t cols, const std::vector<Ciphertext> &e) : dims{ rows, cols }, elems(e) {}
Matrix::Matrix(std::size_t rows, std::size_t cols, std::vector<Ciphertext> &&e) : dims{ rows, cols }, elems(std::move(e)) {}
Matrix::Matrix(std::size_t rows, std::size_t cols) : dims{ rows, cols }, elems(rows * cols) {}

std::vector<std::size_t> Matrix::get_dims() const
{
    return transposed ? std::vector<std::size_t>{ dims[1], dims[0] } : dims;
}
void Matrix::transp() { transposed = !transposed; }
bool Matrix::get_transp() const { return transposed; }
const std::vector<Ciphertext> &Matrix::get_elems() const { return elems; }

// column-major storage; a transposed view swaps the roles of i and j (he_linalg.cpp:376-384)
std::size_t Matrix::ij_to_idx(std::size_t i, std::size_t j) const { return transposed ? j + dims[0] * i : i + dims[0] * j; }
std::size_t Matrix::idx_to_idx(bool colwise, std::size_t idx) const
{
    return transposed != colwise ? idx : idx / dims[1] + idx % dims[1] * dims[0];
}
const Ciphertext &Matrix::operator()(bool colwise, std::size_t idx, bool) const { return elems[idx_to_idx(colwise, idx)]; }
Ciphertext &Matrix::operator()(bool colwise, std::size_t idx, bool) { return elems[idx_to_idx(colwise, idx)]; }
const Ciphertext &Matrix::operator()(std::size_t i, std::size_t j) const { return elems[ij_to_idx(i, j)]; }
Ciphertext &Matrix::operator()(std::size_t i, std::size_t j) { return elems[ij_to_idx(i, j)]; }
void Matrix::set_elem(std::size_t i, std::size_t j, const Ciphertext &e) { elems[ij_to_idx(i, j)] = e; }
void Matrix::set_elem(std::size_t i, std::size_t j, Ciphertext &&e) { elems[ij_to_idx(i, j)] = std::move(e); }

Matrix &Matrix::operator-=(const Evaluator &eval)
{
    for (auto &e : elems) e -= eval;
    return *this;
}
Matrix operator-(const std::tuple<const Evaluator &, const Matrix &> &a)
{
    Matrix res = std::get<1>(a);
    return res -= std::get<0>(a);
}

namespace {
// visit logical entries in the reference's order (j outer, i inner)
template <class F>
void for_each_entry(const std::vector<std::size_t> &d, F &&f)
{
    for (std::size_t j = 0; j < d[1]; ++j)
        for (std::size_t i = 0; i < d[0]; ++i) f(i, j);
}
}  // namespace

Matrix &Matrix::operator+=(const std::tuple<const Evaluator &, const Matrix &> &t)
{
    const Evaluator &eval = std::get<0>(t);
    const Matrix &other = std::get<1>(t);
    assert(get_dims() == other.get_dims());
    for_each_entry(get_dims(), [&](std::size_t i, std::size_t j) { (*this)(i, j) += eval % other(i, j); });
    return *this;
}
Matrix operator+(const std::tuple<const Evaluator &, const Matrix &> &a, const Matrix &op2)
{
    Matrix res = std::get<1>(a);
    return res += std::get<0>(a) % op2;
}
Matrix &Matrix::operator-=(const std::tuple<const Evaluator &, const Matrix &> &t)
{
    const Evaluator &eval = std::get<0>(t);
    const Matrix &other = std::get<1>(t);
    assert(get_dims() == other.get_dims());
    for_each_entry(get_dims(), [&](std::size_t i, std::size_t j) { (*this)(i, j) -= eval % other(i, j); });
    return *this;
}
Matrix operator-(const std::tuple<const Evaluator &, const Matrix &> &a, const Matrix &op2)
{
    Matrix res = std::get<1>(a);
    return res -= std::get<0>(a) % op2;
}
Matrix &Matrix::operator*=(const std::tuple<const EvalRk &, const Matrix &> &t)
{
    const Evaluator &eval = std::get<0>(std::get<0>(t));
    const RelinKeys &rk = std::get<1>(std::get<0>(t));
    const Matrix &other = std::get<1>(t);
    assert(get_dims() == other.get_dims());
    for_each_entry(get_dims(), [&](std::size_t i, std::size_t j) {
        Ciphertext &r = (*this)(i, j);
        r *= eval % other(i, j);
        r &= eval % rk;
        r ^= eval;
    });
    return *this;
}
Matrix operator*(const std::tuple<const EvalRk &, const Matrix &> &a, const Matrix &op2)
{
    Matrix res = std::get<1>(a);
    return res *= std::get<0>(a) % op2;
}

Matrix Matrix::matmul(const Evaluator &eval, const RelinKeys &, const Matrix &other) const
{
    const auto d1 = get_dims(), d2 = other.get_dims();
    assert(d1[1] == d2[0]);
    return device_matmul(eval, elems, transposed, other.elems, other.transposed, d1[0], d1[1], d2[1]);
}
Matrix Matrix::left_matmul_with_transp(const Evaluator &eval, const RelinKeys &) const
{
    const auto d = get_dims();  // A is d[0] x d[1]; result = A^T A is d[1] x d[1]
    return device_matmul(eval, elems, !transposed, elems, transposed, d[1], d[0], d[1]);
}
Matrix Matrix::matmul_square(const Evaluator &eval, const RelinKeys &) const
{
    assert(dims[0] == dims[1]);
    return device_matmul(eval, elems, transposed, elems, transposed, dims[0], dims[0], dims[0]);
}
Matrix Matrix::matmul_pow(const Evaluator &eval, const RelinKeys &rk, int powr) const
{
    assert(dims[0] == dims[1]);
    Matrix res(dims[0], dims[0]);
    Matrix sq = *this;
    bool have = false;
    if (powr & 1) {
        res = sq;
        have = true;
    }
    const int bits = (int)std::ceil(std::log2((double)powr + 1));
    for (int i = 1; i < bits; ++i) {  // square-and-multiply, least significant bit first
        sq = sq.matmul_square(eval, rk);
        if ((powr >> i) & 1) {
            res = have ? res.matmul(eval, rk, sq) : sq;
            have = true;
        }
    }
    return res;
}

// =============================================================================== BatchedVector
BatchedVector::BatchedVector(std::size_t d, const Ciphertext &c) : dim(d), bvec(c) {}
BatchedVector::BatchedVector(std::size_t d, Ciphertext &&c) : dim(d), bvec(std::move(c)) {}
std::size_t BatchedVector::get_dim() const { return dim; }
const Ciphertext &BatchedVector::get_bvec() const { return bvec; }

BatchedVector &BatchedVector::operator-=(const Evaluator &eval)
{
    bvec -= eval;
    return *this;
}
BatchedVector operator-(const std::tuple<const Evaluator &, const BatchedVector &> &a)
{
    BatchedVector res = std::get<1>(a);
    return res -= std::get<0>(a);
}
BatchedVector &BatchedVector::operator+=(const std::tuple<const Evaluator &, const BatchedVector &> &t)
{
    bvec += std::get<0>(t) % std::get<1>(t).bvec;
    return *this;
}
BatchedVector operator+(const std::tuple<const Evaluator &, const BatchedVector &> &a, const BatchedVector &op2)
{
    BatchedVector res = std::get<1>(a);
    return res += std::get<0>(a) % op2;
}
BatchedVector &BatchedVector::operator-=(const std::tuple<const Evaluator &, const BatchedVector &> &t)
{
    bvec -= std::get<0>(t) % std::get<1>(t).bvec;
    return *this;
}
BatchedVector operator-(const std::tuple<const Evaluator &, const BatchedVector &> &a, const BatchedVector &op2)
{
    BatchedVector res = std::get<1>(a);
    return res -= std::get<0>(a) % op2;
}
BatchedVector &BatchedVector::operator*=(const std::tuple<const Evaluator &, const BatchedVector &> &t)
{
    bvec *= std::get<0>(t) % std::get<1>(t).bvec;
    return *this;
}
BatchedVector operator*(const std::tuple<const Evaluator &, const BatchedVector &> &a, const BatchedVector &op2)
{
    BatchedVector res = std::get<1>(a);
    return res *= std::get<0>(a) % op2;
}
BatchedVector &BatchedVector::operator&=(const EvalRk &k)
{
    bvec &= std::get<0>(k) % std::get<1>(k);
    return *this;
}
BatchedVector operator&(const EvalRk &k, const BatchedVector &op)
{
    BatchedVector res = op;
    return res &= k;
}
BatchedVector &BatchedVector::operator^=(const Evaluator &eval)
{
    bvec ^= eval;
    return *this;
}
BatchedVector operator^(const Evaluator &eval, const BatchedVector &op)
{
    BatchedVector res = op;
    return res ^= eval;
}
BatchedVector &BatchedVector::operator*=(const std::tuple<const EvalRk &, const BatchedVector &> &t)
{
    const Evaluator &eval = std::get<0>(std::get<0>(t));
    bvec *= eval % std::get<1>(t).bvec;
    bvec &= eval % std::get<1>(std::get<0>(t));
    bvec ^= eval;
    return *this;
}
BatchedVector operator*(const std::tuple<const EvalRk &, const BatchedVector &> &a, const BatchedVector &op2)
{
    BatchedVector res = std::get<1>(a);
    return res *= std::get<0>(a) % op2;
}
BatchedVector &BatchedVector::operator<<=(const std::tuple<const EvalGk &, const int &> &t)
{
    const EvalGk &k = std::get<0>(t);
    std::get<0>(k).rotate_vector_inplace(bvec, std::get<1>(t), std::get<1>(k));
    return *this;
}
BatchedVector operator<<(const std::tuple<const EvalGk &, const BatchedVector &> &a, int steps)
{
    const EvalGk &k = std::get<0>(a);
    BatchedVector res = std::get<1>(a);
    std::get<0>(k).rotate_vector(std::get<1>(a).bvec, steps, std::get<1>(k), res.bvec);
    return res;
}
BatchedVector &BatchedVector::operator>>=(const std::tuple<const EvalGk &, const int &> &t)
{
    const EvalGk &k = std::get<0>(t);
    std::get<0>(k).rotate_vector_inplace(bvec, -std::get<1>(t), std::get<1>(k));
    return *this;
}
BatchedVector operator>>(const std::tuple<const EvalGk &, const BatchedVector &> &a, int steps)
{
    const EvalGk &k = std::get<0>(a);
    BatchedVector res = std::get<1>(a);
    std::get<0>(k).rotate_vector(std::get<1>(a).bvec, -steps, std::get<1>(k), res.bvec);
    return res;
}

BatchedVector &BatchedVector::square_inplace(const Evaluator &eval, const RelinKeys &rk)
{
    eval.square_inplace(bvec);
    bvec &= eval % rk;
    bvec ^= eval;
    return *this;
}
BatchedVector BatchedVector::square(const Evaluator &eval, const RelinKeys &rk) const
{
    BatchedVector res = *this;
    res.square_inplace(eval, rk);
    return res;
}

// Rotate-and-add reduction of the first `dim` slots into slot 0 (he_linalg.cpp:667-713): dim is
// read as a binary number; every set bit 2^b contributes a log-depth rotate/add tree over a
// window of 2^b slots, and the remaining input is shifted past each window that was consumed.
BatchedVector &BatchedVector::sum_elems_inplace(const Evaluator &eval, const GaloisKeys &gk)
{
    Ciphertext rest = bvec;  // slots not yet summed, shifted so that the next window starts at slot 0
    Ciphertext window_sum, rotated;
    bool have_total = false;  // bvec already holds a partial total
    int window = 1;
    if (dim & 1) {  // window of one slot: bvec itself is the partial total
        have_total = true;
        rest <<= eval % gk % window;
    }
    for (std::size_t bits = dim >> 1; bits != 0; bits >>= 1) {
        window <<= 1;
        if (!(bits & 1)) continue;
        int steps = window >> 1;
        rotated = eval % gk % rest << steps;
        Ciphertext &acc = have_total ? window_sum : bvec;
        acc = eval % rest + rotated;
        for (steps >>= 1; steps != 0; steps >>= 1) {
            rotated = eval % gk % acc << steps;
            acc += eval % rotated;
        }
        if (have_total)
            bvec += eval % window_sum;
        else
            have_total = true;
        if (bits != 1) rest <<= eval % gk % window;
    }
    dim = 1;
    return *this;
}
BatchedVector BatchedVector::sum_elems(const Evaluator &eval, const GaloisKeys &gk)
{
    BatchedVector res = *this;
    res.sum_elems_inplace(eval, gk);
    return res;
}

// =============================================================================== BatchedMatrix
BatchedMatrix::BatchedMatrix(BatchingType t, const std::vector<BatchedVector> &v) : btype(t), bvecs(v) {}
BatchedMatrix::BatchedMatrix(BatchingType t, std::vector<BatchedVector> &&v) : btype(t), bvecs(std::move(v)) {}
BatchedMatrix::BatchingType BatchedMatrix::get_btype() const { return btype; }
bool BatchedMatrix::get_transp() const { return transposed; }
void BatchedMatrix::transp() { transposed = !transposed; }
std::size_t BatchedMatrix::get_col_dim() const { return !transposed ? bvecs.size() : bvecs[0].get_dim(); }
std::size_t BatchedMatrix::get_row_dim() const { return transposed ? bvecs.size() : bvecs[0].get_dim(); }
const std::vector<BatchedVector> &BatchedMatrix::get_bvecs() const { return bvecs; }
const BatchedVector &BatchedMatrix::operator[](std::size_t i) const { return bvecs[i]; }
BatchedVector &BatchedMatrix::operator[](std::size_t i) { return bvecs[i]; }

BatchedMatrix &BatchedMatrix::operator-=(const Evaluator &eval)
{
    for (auto &v : bvecs) v -= eval;
    return *this;
}
BatchedMatrix operator-(const std::tuple<const Evaluator &, const BatchedMatrix &> &a)
{
    BatchedMatrix res = std::get<1>(a);
    return res -= std::get<0>(a);
}
BatchedMatrix &BatchedMatrix::operator+=(const std::tuple<const Evaluator &, const BatchedMatrix &> &t)
{
    const BatchedMatrix &other = std::get<1>(t);
    assert(transposed == other.transposed);
    for (std::size_t i = 0; i < bvecs.size(); ++i) bvecs[i] += std::get<0>(t) % other.bvecs[i];
    return *this;
}
BatchedMatrix operator+(const std::tuple<const Evaluator &, const BatchedMatrix &> &a, const BatchedMatrix &op2)
{
    BatchedMatrix res = std::get<1>(a);
    return res += std::get<0>(a) % op2;
}
BatchedMatrix &BatchedMatrix::operator-=(const std::tuple<const Evaluator &, const BatchedMatrix &> &t)
{
    const BatchedMatrix &other = std::get<1>(t);
    assert(transposed == other.transposed);
    for (std::size_t i = 0; i < bvecs.size(); ++i) bvecs[i] -= std::get<0>(t) % other.bvecs[i];
    return *this;
}
BatchedMatrix operator-(const std::tuple<const Evaluator &, const BatchedMatrix &> &a, const BatchedMatrix &op2)
{
    BatchedMatrix res = std::get<1>(a);
    return res -= std::get<0>(a) % op2;
}
BatchedMatrix &BatchedMatrix::operator*=(const std::tuple<const EvalRk &, const BatchedMatrix &> &t)
{
    const BatchedMatrix &other = std::get<1>(t);
    assert(transposed == other.transposed);
    for (std::size_t i = 0; i < bvecs.size(); ++i) bvecs[i] *= std::get<0>(t) % other.bvecs[i];
    return *this;
}
BatchedMatrix operator*(const std::tuple<const EvalRk &, const BatchedMatrix &> &a, const BatchedMatrix &op2)
{
    BatchedMatrix res = std::get<1>(a);
    return res *= std::get<0>(a) % op2;
}
BatchedMatrix &BatchedMatrix::square_inplace(const Evaluator &eval, const RelinKeys &rk)
{
    for (auto &v : bvecs) v.square_inplace(eval, rk);
    return *this;
}
BatchedMatrix BatchedMatrix::square(const Evaluator &eval, const RelinKeys &rk) const
{
    BatchedMatrix res = *this;
    res.square_inplace(eval, rk);
    return res;
}
BatchedMatrix &BatchedMatrix::sum_bvec_elems_inplace(const Evaluator &eval, const GaloisKeys &gk)
{
    for (auto &v : bvecs) v.sum_elems_inplace(eval, gk);
    return *this;
}
BatchedMatrix BatchedMatrix::sum_bvec_elems(const Evaluator &eval, const GaloisKeys &gk)
{
    BatchedMatrix res = *this;
    res.sum_bvec_elems_inplace(eval, gk);
    return res;
}

// he_linalg.cpp:943-1006.  Case A: this = n generalised diagonals, other = p columns,
// res_i = sum_j rot(other_i, j) * this_j (column batching).  Case B: this = n columns, other =
// n transposed columns (rows of length p), res_i = sum_j rot(other_j, i) * this_j (diagonal
// batching).  One relinearisation and one rescale per output (SMART_RELIN, :975).
BatchedMatrix BatchedMatrix::matmul(const Evaluator &eval, const RelinKeys &, const GaloisKeys &, const BatchedMatrix &other) const
{
    assert(other.btype == BatchingType::col);
    assert(!transposed);
    const bool case_b = btype == BatchingType::col;
    const std::size_t n = get_col_dim(), p = other.get_col_dim();
    if (case_b) {
        assert(other.transposed);
        assert(n == other.get_row_dim());
    } else {
        assert(!other.transposed);
    }
    const SEALContext &ctx = eval.context();
    Batch bt, bo, bout;
    gather(ctx, bt, bvecs.size(), 2, [&](std::size_t i) -> const Ciphertext & { return bvecs[i].get_bvec(); });
    gather(ctx, bo, other.bvecs.size(), 2, [&](std::size_t i) -> const Ciphertext & { return other.bvecs[i].get_bvec(); });
    check(hegpu_ct_create(ctx.raw(), &bout.h, (std::uint32_t)p, 2, ctx.key_limbs() - 1));
    check(hegpu_bmatmul(ctx.raw(), bout.h, bt.h, bo.h, (std::uint32_t)n, (std::uint32_t)p, case_b ? 1 : 0));
    std::vector<Ciphertext> out(p);
    scatter(ctx, bout, p, [&](std::size_t i) -> Ciphertext & { return out[i]; });
    std::vector<BatchedVector> res;
    res.reserve(p);
    // result vectors inherit the logical dimension of the rotated operand (`other`)
    for (std::size_t i = 0; i < p; ++i) res.emplace_back(other.bvecs[case_b ? 0 : i].get_dim(), std::move(out[i]));
    return BatchedMatrix(case_b ? BatchingType::diag : BatchingType::col, std::move(res));
}

}  // namespace he::linalg
