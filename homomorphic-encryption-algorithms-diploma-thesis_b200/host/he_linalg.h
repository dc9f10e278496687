// he_linalg.h -- host mirror of include/he_linalg.h:47-412: Matrix (one ciphertext per entry,
// column-major, transposition by flag), BatchedVector (one slot-batched ciphertext + logical
// dimension) and BatchedMatrix (columns or generalised diagonals as BatchedVectors), with
// the reference's method and operator names over he::gpu types.
//
// The element-wise operators forward to he::operators one evaluator call at a time, as the
// reference does.  The three routines on the hot path do not: Matrix::matmul,
// BatchedMatrix::matmul and the *_square / matmul_pow family gather their operands into
// device batches and run ONE composite of the C ABI (hegpu_matmul_elemwise / hegpu_bmatmul),
// which replays the reference's loop order bit for bit.
#pragma once
#include <cstddef>
#include <tuple>
#include <vector>

#include "he_operators.h"

namespace he::linalg {

using he::gpu::Ciphertext;
using he::gpu::Evaluator;
using he::gpu::GaloisKeys;
using he::gpu::RelinKeys;

class Matrix;
class BatchedVector;
class BatchedMatrix;

// the ciphertext-level operators stay visible next to the overloads declared here
using he::operators::operator%;
using he::operators::operator+;
using he::operators::operator-;
using he::operators::operator*;
using he::operators::operator&;
using he::operators::operator^;
using he::operators::operator|;
using he::operators::operator<<;
using he::operators::operator>>;

// ties: eval % M, eval % rk % M, eval % gk % M (include/he_linalg.h:13-43)
template <typename T>
concept LinalgOperand_tn = std::same_as<std::decay_t<T>, Matrix> || std::same_as<std::decay_t<T>, BatchedVector> ||
                           std::same_as<std::decay_t<T>, BatchedMatrix>;
template <LinalgOperand_tn T>
constexpr auto operator%(const Evaluator &eval, T &&op)
{
    return std::tie(eval, std::forward<T>(op));
}
template <LinalgOperand_tn T>
constexpr auto operator%(const std::tuple<const Evaluator &, const RelinKeys &> &eval_rk, T &&op)
{
    return std::tie(eval_rk, std::forward<T>(op));
}
template <LinalgOperand_tn T>
constexpr auto operator%(const std::tuple<const Evaluator &, const GaloisKeys &> &eval_gk, T &&op)
{
    return std::tie(eval_gk, std::forward<T>(op));
}

using EvalRk = std::tuple<const Evaluator &, const RelinKeys &>;
using EvalGk = std::tuple<const Evaluator &, const GaloisKeys &>;

// ------------------------------------------------------------------------------- Matrix
class Matrix {
public:
    Matrix() = delete;
    Matrix(std::size_t rows, std::size_t cols, const std::vector<Ciphertext> &elems);
    Matrix(std::size_t rows, std::size_t cols, std::vector<Ciphertext> &&elems);
    Matrix(std::size_t rows, std::size_t cols);

    std::vector<std::size_t> get_dims() const;
    void transp();
    bool get_transp() const;
    const std::vector<Ciphertext> &get_elems() const;
    const Ciphertext &operator()(bool colwise, std::size_t idx, bool dummy_arg) const;
    Ciphertext &operator()(bool colwise, std::size_t idx, bool dummy_arg);
    const Ciphertext &operator()(std::size_t i, std::size_t j) const;
    Ciphertext &operator()(std::size_t i, std::size_t j);
    void set_elem(std::size_t i, std::size_t j, const Ciphertext &elem);
    void set_elem(std::size_t i, std::size_t j, Ciphertext &&elem);

    Matrix &operator-=(const Evaluator &eval);                                   // negate
    friend Matrix operator-(const std::tuple<const Evaluator &, const Matrix &> &eval_op);
    Matrix &operator+=(const std::tuple<const Evaluator &, const Matrix &> &eval_other);  // element-wise
    friend Matrix operator+(const std::tuple<const Evaluator &, const Matrix &> &eval_op1, const Matrix &op2);
    Matrix &operator-=(const std::tuple<const Evaluator &, const Matrix &> &eval_other);
    friend Matrix operator-(const std::tuple<const Evaluator &, const Matrix &> &eval_op1, const Matrix &op2);
    // element-wise multiply with relinearize + rescale (he_linalg.cpp:158-182)
    Matrix &operator*=(const std::tuple<const EvalRk &, const Matrix &> &eval_rk__other);
    friend Matrix operator*(const std::tuple<const EvalRk &, const Matrix &> &eval_rk__op1, const Matrix &op2);

    Matrix matmul(const Evaluator &eval, const RelinKeys &rk, const Matrix &other) const;
    Matrix left_matmul_with_transp(const Evaluator &eval, const RelinKeys &rk) const;  // A^T A
    Matrix matmul_square(const Evaluator &eval, const RelinKeys &rk) const;            // A A
    Matrix matmul_pow(const Evaluator &eval, const RelinKeys &rk, int powr) const;     // square-and-multiply

private:
    std::size_t ij_to_idx(std::size_t i, std::size_t j) const;
    std::size_t idx_to_idx(bool colwise, std::size_t idx) const;
    std::vector<std::size_t> dims;
    bool transposed = false;
    std::vector<Ciphertext> elems;
};

// ------------------------------------------------------------------------------- BatchedVector
class BatchedVector {
public:
    BatchedVector() = delete;
    BatchedVector(std::size_t dim, const Ciphertext &bvec);
    BatchedVector(std::size_t dim, Ciphertext &&bvec);
    std::size_t get_dim() const;
    const Ciphertext &get_bvec() const;

    BatchedVector &operator-=(const Evaluator &eval);
    friend BatchedVector operator-(const std::tuple<const Evaluator &, const BatchedVector &> &eval_op);
    BatchedVector &operator+=(const std::tuple<const Evaluator &, const BatchedVector &> &eval_other);
    friend BatchedVector operator+(const std::tuple<const Evaluator &, const BatchedVector &> &eval_op1, const BatchedVector &op2);
    BatchedVector &operator-=(const std::tuple<const Evaluator &, const BatchedVector &> &eval_other);
    friend BatchedVector operator-(const std::tuple<const Evaluator &, const BatchedVector &> &eval_op1, const BatchedVector &op2);
    BatchedVector &operator*=(const std::tuple<const Evaluator &, const BatchedVector &> &eval_other);  // no relin
    friend BatchedVector operator*(const std::tuple<const Evaluator &, const BatchedVector &> &eval_op1, const BatchedVector &op2);
    BatchedVector &operator&=(const EvalRk &eval_rk);  // relinearize
    friend BatchedVector operator&(const EvalRk &eval_rk, const BatchedVector &op);
    BatchedVector &operator^=(const Evaluator &eval);  // rescale
    friend BatchedVector operator^(const Evaluator &eval, const BatchedVector &op);
    BatchedVector &operator*=(const std::tuple<const EvalRk &, const BatchedVector &> &eval_rk__other);  // multiply, relin, rescale
    friend BatchedVector operator*(const std::tuple<const EvalRk &, const BatchedVector &> &eval_rk__op1, const BatchedVector &op2);
    BatchedVector &operator<<=(const std::tuple<const EvalGk &, const int &> &eval_gk__steps);
    friend BatchedVector operator<<(const std::tuple<const EvalGk &, const BatchedVector &> &eval_gk__op, int steps);
    BatchedVector &operator>>=(const std::tuple<const EvalGk &, const int &> &eval_gk__steps);
    friend BatchedVector operator>>(const std::tuple<const EvalGk &, const BatchedVector &> &eval_gk__op, int steps);

    BatchedVector &square_inplace(const Evaluator &eval, const RelinKeys &rk);
    BatchedVector square(const Evaluator &eval, const RelinKeys &rk) const;
    BatchedVector &sum_elems_inplace(const Evaluator &eval, const GaloisKeys &gk);
    BatchedVector sum_elems(const Evaluator &eval, const GaloisKeys &gk);

private:
    std::size_t dim;
    Ciphertext bvec;
};

// ------------------------------------------------------------------------------- BatchedMatrix
class BatchedMatrix {
public:
    enum class BatchingType { col, diag };
    BatchedMatrix() = delete;
    BatchedMatrix(BatchingType btype, const std::vector<BatchedVector> &bvecs);
    BatchedMatrix(BatchingType btype, std::vector<BatchedVector> &&bvecs);

    BatchingType get_btype() const;
    std::size_t get_col_dim() const;
    std::size_t get_row_dim() const;
    bool get_transp() const;
    void transp();
    const std::vector<BatchedVector> &get_bvecs() const;
    const BatchedVector &operator[](std::size_t i) const;
    BatchedVector &operator[](std::size_t i);

    BatchedMatrix &operator-=(const Evaluator &eval);
    friend BatchedMatrix operator-(const std::tuple<const Evaluator &, const BatchedMatrix &> &eval_op);
    BatchedMatrix &operator+=(const std::tuple<const Evaluator &, const BatchedMatrix &> &eval_other);
    friend BatchedMatrix operator+(const std::tuple<const Evaluator &, const BatchedMatrix &> &eval_op1, const BatchedMatrix &op2);
    BatchedMatrix &operator-=(const std::tuple<const Evaluator &, const BatchedMatrix &> &eval_other);
    friend BatchedMatrix operator-(const std::tuple<const Evaluator &, const BatchedMatrix &> &eval_op1, const BatchedMatrix &op2);
    BatchedMatrix &operator*=(const std::tuple<const EvalRk &, const BatchedMatrix &> &eval_rk__other);  // element-wise
    friend BatchedMatrix operator*(const std::tuple<const EvalRk &, const BatchedMatrix &> &eval_rk__op1, const BatchedMatrix &op2);

    BatchedMatrix &square_inplace(const Evaluator &eval, const RelinKeys &rk);
    BatchedMatrix square(const Evaluator &eval, const RelinKeys &rk) const;
    BatchedMatrix &sum_bvec_elems_inplace(const Evaluator &eval, const GaloisKeys &gk);
    BatchedMatrix sum_bvec_elems(const Evaluator &eval, const GaloisKeys &gk);

    // he_linalg.cpp:943-1006: case A (this = diag, other = col) or case B (this = col, other = transposed col)
    BatchedMatrix matmul(const Evaluator &eval, const RelinKeys &rk, const GaloisKeys &gk, const BatchedMatrix &other) const;

private:
    BatchingType btype;
    bool transposed = false;
    std::vector<BatchedVector> bvecs;
};

}  // namespace he::linalg
