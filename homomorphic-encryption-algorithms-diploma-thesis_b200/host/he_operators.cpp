// he_operators.cpp -- the 26 operator overloads, each forwarding to exactly one evaluator
// method (mirror of src/core/he_operators.cpp:14-237; the in-place form calls X_inplace, the
// value form calls the out-of-place method into a fresh result).
#include "he_operators.h"

namespace he::operators {

namespace {
// value forms share one shape: build a result with an out-of-place evaluator call
template <class F>
inline Ciphertext made(F &&call)
{
    Ciphertext res;
    call(res);
    return res;
}
inline const Evaluator &ev(const EvalCt &t) { return std::get<0>(t); }
inline const Ciphertext &ct(const EvalCt &t) { return std::get<1>(t); }
}  // namespace

// negate
Ciphertext &operator-=(Ciphertext &op, const Evaluator &eval) { return eval.negate_inplace(op), op; }
Ciphertext operator-(const EvalCt &a) { return made([&](Ciphertext &r) { ev(a).negate(ct(a), r); }); }

// add / add_plain
Ciphertext &operator+=(Ciphertext &op1, const EvalCt &b) { return ev(b).add_inplace(op1, ct(b)), op1; }
Ciphertext operator+(const EvalCt &a, const Ciphertext &op2) { return made([&](Ciphertext &r) { ev(a).add(ct(a), op2, r); }); }
Ciphertext &operator+=(Ciphertext &op1, const EvalPt &b) { return std::get<0>(b).add_plain_inplace(op1, std::get<1>(b)), op1; }
Ciphertext operator+(const EvalCt &a, const Plaintext &op2) { return made([&](Ciphertext &r) { ev(a).add_plain(ct(a), op2, r); }); }

// sub / sub_plain
Ciphertext &operator-=(Ciphertext &op1, const EvalCt &b) { return ev(b).sub_inplace(op1, ct(b)), op1; }
Ciphertext operator-(const EvalCt &a, const Ciphertext &op2) { return made([&](Ciphertext &r) { ev(a).sub(ct(a), op2, r); }); }
Ciphertext &operator-=(Ciphertext &op1, const EvalPt &b) { return std::get<0>(b).sub_plain_inplace(op1, std::get<1>(b)), op1; }
Ciphertext operator-(const EvalCt &a, const Plaintext &op2) { return made([&](Ciphertext &r) { ev(a).sub_plain(ct(a), op2, r); }); }

// multiply / multiply_plain
Ciphertext &operator*=(Ciphertext &op1, const EvalCt &b) { return ev(b).multiply_inplace(op1, ct(b)), op1; }
Ciphertext operator*(const EvalCt &a, const Ciphertext &op2) { return made([&](Ciphertext &r) { ev(a).multiply(ct(a), op2, r); }); }
Ciphertext &operator*=(Ciphertext &op1, const EvalPt &b) { return std::get<0>(b).multiply_plain_inplace(op1, std::get<1>(b)), op1; }
Ciphertext operator*(const EvalCt &a, const Plaintext &op2) { return made([&](Ciphertext &r) { ev(a).multiply_plain(ct(a), op2, r); }); }

// relinearize
Ciphertext &operator&=(Ciphertext &op, const EvalRk &k) { return std::get<0>(k).relinearize_inplace(op, std::get<1>(k)), op; }
Ciphertext operator&(const EvalRk &k, const Ciphertext &op)
{
    return made([&](Ciphertext &r) { std::get<0>(k).relinearize(op, std::get<1>(k), r); });
}

// rescale, mod switch
Ciphertext &operator^=(Ciphertext &op, const Evaluator &eval) { return eval.rescale_to_next_inplace(op), op; }
Ciphertext operator^(const Evaluator &eval, const Ciphertext &op) { return made([&](Ciphertext &r) { eval.rescale_to_next(op, r); }); }
Ciphertext &operator|=(Ciphertext &op, const Evaluator &eval) { return eval.mod_switch_to_next_inplace(op), op; }
Ciphertext operator|(const Evaluator &eval, const Ciphertext &op) { return made([&](Ciphertext &r) { eval.mod_switch_to_next(op, r); }); }

// rotations: << left, >> right (= left by -steps)
Ciphertext &operator<<=(Ciphertext &op, const EvalGkInt &a)
{
    const EvalGk &k = std::get<0>(a);
    return std::get<0>(k).rotate_vector_inplace(op, std::get<1>(a), std::get<1>(k)), op;
}
Ciphertext operator<<(const EvalGkCt &a, int steps)
{
    const EvalGk &k = std::get<0>(a);
    return made([&](Ciphertext &r) { std::get<0>(k).rotate_vector(std::get<1>(a), steps, std::get<1>(k), r); });
}
Ciphertext &operator>>=(Ciphertext &op, const EvalGkInt &a)
{
    const EvalGk &k = std::get<0>(a);
    return std::get<0>(k).rotate_vector_inplace(op, -std::get<1>(a), std::get<1>(k)), op;
}
Ciphertext operator>>(const EvalGkCt &a, int steps)
{
    const EvalGk &k = std::get<0>(a);
    return made([&](Ciphertext &r) { std::get<0>(k).rotate_vector(std::get<1>(a), -steps, std::get<1>(k), r); });
}

}  // namespace he::operators
