// he_operators.h -- host mirror of the reference's operator DSL (include/he_operators.h:1-160):
// the same 26 overloads on the same tuple ties, over he::gpu value types.  Every operator is
// exactly one evaluator call, i.e. one C-ABI entry point (src/core/he_operators.cpp:14-237).
//
//   eval % ct                      tie an evaluator with an operand
//   eval % gk % ct                 tie for rotations
//   ct -= eval          -(eval % ct)                 negate
//   a += eval % b       (eval % a) + b               add           (ciphertext or plaintext b)
//   a -= eval % b       (eval % a) - b               sub           (ciphertext or plaintext b)
//   a *= eval % b       (eval % a) * b               multiply      (ciphertext or plaintext b)
//   a &= eval % rk      (eval % rk) & a              relinearize
//   a ^= eval           ^ is rescale_to_next;        a |= eval     mod_switch_to_next
//   a <<= eval % gk % k (eval % gk % a) << k         rotate left;  >>= / >> rotate right
#pragma once
#include <tuple>
#include <utility>

#include "hegpu_seal_like.hpp"

namespace he::operators {

using he::gpu::Ciphertext;
using he::gpu::Evaluator;
using he::gpu::GaloisKeys;
using he::gpu::Plaintext;
using he::gpu::RelinKeys;

template <typename T>
concept Operand_tn = std::same_as<std::decay_t<T>, Plaintext> || std::same_as<std::decay_t<T>, Ciphertext> ||
                     std::same_as<std::decay_t<T>, GaloisKeys> || std::same_as<std::decay_t<T>, RelinKeys>;

template <Operand_tn T>
constexpr auto operator%(const Evaluator &eval, T &&op)
{
    return std::tie(eval, std::forward<T>(op));
}

template <typename T>
concept Ciphertext_int_tn = std::same_as<std::decay_t<T>, Ciphertext> || std::same_as<std::decay_t<T>, int>;

template <Ciphertext_int_tn T>
constexpr auto operator%(const std::tuple<const Evaluator &, const GaloisKeys &> &eval_gk, T &&op)
{
    return std::forward_as_tuple(eval_gk, std::forward<T>(op));  // also accepts a literal step count
}

using EvalCt = std::tuple<const Evaluator &, const Ciphertext &>;
using EvalPt = std::tuple<const Evaluator &, const Plaintext &>;
using EvalRk = std::tuple<const Evaluator &, const RelinKeys &>;
using EvalGk = std::tuple<const Evaluator &, const GaloisKeys &>;
using EvalGkCt = std::tuple<const EvalGk &, const Ciphertext &>;
using EvalGkInt = std::tuple<const EvalGk &, const int &>;

Ciphertext &operator-=(Ciphertext &op, const Evaluator &eval);
Ciphertext operator-(const EvalCt &eval_op);

Ciphertext &operator+=(Ciphertext &op1, const EvalCt &eval_op2);
Ciphertext operator+(const EvalCt &eval_op1, const Ciphertext &op2);
Ciphertext &operator+=(Ciphertext &op1, const EvalPt &eval_op2);
Ciphertext operator+(const EvalCt &eval_op1, const Plaintext &op2);

Ciphertext &operator-=(Ciphertext &op1, const EvalCt &eval_op2);
Ciphertext operator-(const EvalCt &eval_op1, const Ciphertext &op2);
Ciphertext &operator-=(Ciphertext &op1, const EvalPt &eval_op2);
Ciphertext operator-(const EvalCt &eval_op1, const Plaintext &op2);

Ciphertext &operator*=(Ciphertext &op1, const EvalCt &eval_op2);
Ciphertext operator*(const EvalCt &eval_op1, const Ciphertext &op2);
Ciphertext &operator*=(Ciphertext &op1, const EvalPt &eval_op2);
Ciphertext operator*(const EvalCt &eval_op1, const Plaintext &op2);

Ciphertext &operator&=(Ciphertext &op, const EvalRk &eval_rk);
Ciphertext operator&(const EvalRk &eval_rk, const Ciphertext &op);

Ciphertext &operator^=(Ciphertext &op, const Evaluator &eval);
Ciphertext operator^(const Evaluator &eval, const Ciphertext &op);

Ciphertext &operator|=(Ciphertext &op, const Evaluator &eval);
Ciphertext operator|(const Evaluator &eval, const Ciphertext &op);

Ciphertext &operator<<=(Ciphertext &op, const EvalGkInt &eval_gk__steps);
Ciphertext operator<<(const EvalGkCt &eval_gk__op, int steps);
Ciphertext &operator>>=(Ciphertext &op, const EvalGkInt &eval_gk__steps);
Ciphertext operator>>(const EvalGkCt &eval_gk__op, int steps);

}  // namespace he::operators
