// he_server.hpp -- the reference server's modes (src/demos/server.cpp) over byte buffers: each function takes the
// payload the reference's server reads after the streampos size prefix (SEAL-serialized parameters, relinearisation
// keys and ciphertexts, exactly what src/demos/client.cpp sends) and returns the payload it writes back, with the
// computation on the GPU evaluator.  Socket setup (server.cpp:27-90) is I/O and stays out of scope: a
// caller reads the size prefix and the dimensions, hands the buffer over and writes the reply.
//   server_side_simple        server.cpp:92-159    multiply, relinearize, rescale
//   server_side_batch_matmul  server.cpp:161-244   Matrix::matmul
//   server_side_fft           server.cpp:527-592   he::fft::fft over a vector of ciphertexts
#pragma once
#include <cstddef>

#include "seal_wire.hpp"

namespace he::server {

he::wire::bytes server_side_simple(const he::wire::bytes &request, int device = 0);
he::wire::bytes server_side_batch_matmul(const he::wire::bytes &request, std::size_t mat1_rows, std::size_t mat1_cols, std::size_t mat2_rows,
                                         std::size_t mat2_cols, int device = 0);
he::wire::bytes server_side_fft(const he::wire::bytes &request, std::size_t vec_elem_no, int device = 0);

}  // namespace he::server
