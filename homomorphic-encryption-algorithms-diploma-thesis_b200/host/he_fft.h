// he_fft.h -- host mirror of include/he_fft.h:12-27 (same four entry points, same argument
// order) over he::gpu types.  fft/ifft act on a vector of ciphertexts (one coefficient per
// ciphertext, he_fft.cpp:13-87); bfft/ibfft act on the slots of one ciphertext
// (he_fft.cpp:166-223) and leave the result in bit-reversed order.
#pragma once
#include <complex>
#include <vector>

#include "hegpu_seal_like.hpp"

namespace he::fft {

using he::gpu::Ciphertext;
using he::gpu::CKKSEncoder;
using he::gpu::Evaluator;
using he::gpu::GaloisKeys;

std::vector<Ciphertext> fft(const CKKSEncoder &cencd, const Evaluator &eval, const std::vector<Ciphertext> &vec_ct);
std::vector<Ciphertext> ifft(const CKKSEncoder &cencd, const Evaluator &eval, const std::vector<Ciphertext> &vec_ct);
Ciphertext bfft(const CKKSEncoder &cencd, const Evaluator &eval, const GaloisKeys &gk, const Ciphertext &x_ct, std::size_t n);
Ciphertext ibfft(const CKKSEncoder &cencd, const Evaluator &eval, const GaloisKeys &gk, const Ciphertext &x_ct, std::size_t n);

// test hook: every plaintext the routines encode is appended here when non-null, so that a
// checker can replay the same evaluator calls with bit-identical plaintext limbs
extern std::vector<he::gpu::Plaintext> *encoded_log;

}  // namespace he::fft
