// he_util.h -- host mirror of include/he_util.h:13-77: chain-index lookup and level burning
// (multiply by an encoding of 1 at the ciphertext's scale, then rescale).
#pragma once
#include <vector>

#include "hegpu_seal_like.hpp"

namespace he::util {

using he::gpu::Ciphertext;
using he::gpu::CKKSEncoder;
using he::gpu::Evaluator;
using he::gpu::Plaintext;
using he::gpu::SEALContext;

inline std::size_t get_chain_index(const SEALContext &ctx, const Ciphertext &ct) { return ctx.chain_index(ct.parms_id()); }

// one ciphertext or a set of ciphertexts that sit at the same level
inline void drop_chain_levels(const SEALContext &, const CKKSEncoder &cencd, const Evaluator &eval, Plaintext &one_pt,
                              const std::vector<Ciphertext *> &cts, std::size_t num_of_levels)
{
    for (std::size_t lvl = 0; lvl < num_of_levels; ++lvl) {
        cencd.encode(1.0, cts.front()->parms_id(), cts.front()->scale(), one_pt);
        for (Ciphertext *c : cts) {
            eval.multiply_plain_inplace(*c, one_pt);
            eval.rescale_to_next_inplace(*c);
        }
    }
}
inline void drop_chain_levels(const SEALContext &ctx, const CKKSEncoder &cencd, const Evaluator &eval, Plaintext &one_pt, Ciphertext &ct,
                              std::size_t num_of_levels)
{
    drop_chain_levels(ctx, cencd, eval, one_pt, std::vector<Ciphertext *>{ &ct }, num_of_levels);
}
template <class T>
inline void drop_chain_levels(const SEALContext &ctx, const CKKSEncoder &cencd, const Evaluator &eval, T &&cts, std::size_t num_of_levels)
{
    Plaintext one_pt;
    drop_chain_levels(ctx, cencd, eval, one_pt, std::forward<T>(cts), num_of_levels);
}

inline void reach_chain_level(const SEALContext &ctx, const CKKSEncoder &cencd, const Evaluator &eval, Plaintext &one_pt,
                              const std::vector<Ciphertext *> &cts, const Ciphertext &to_reach_ct)
{
    const std::size_t levels = get_chain_index(ctx, *cts.front()) - get_chain_index(ctx, to_reach_ct);
    drop_chain_levels(ctx, cencd, eval, one_pt, cts, levels);
}
inline void reach_chain_level(const SEALContext &ctx, const CKKSEncoder &cencd, const Evaluator &eval, Plaintext &one_pt, Ciphertext &ct,
                              const Ciphertext &to_reach_ct)
{
    reach_chain_level(ctx, cencd, eval, one_pt, std::vector<Ciphertext *>{ &ct }, to_reach_ct);
}
template <class T>
inline void reach_chain_level(const SEALContext &ctx, const CKKSEncoder &cencd, const Evaluator &eval, T &&cts, const Ciphertext &to_reach_ct)
{
    Plaintext one_pt;
    reach_chain_level(ctx, cencd, eval, one_pt, std::forward<T>(cts), to_reach_ct);
}

}  // namespace he::util
