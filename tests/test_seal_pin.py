"""Real-SEAL pin of the oracle and of the GPU evaluator ("parity unpinned").

tools/seal_golden.cpp, built against Microsoft SEAL 4.1 by whoever has it, dumps keys, input ciphertexts and the
outputs of every evaluator primitive on the path into a HEGVEC1 file.  Drop the file(s) at
tests/golden/seal_vectors*.hegvec (or point HEGPU_SEAL_VECTORS at one) and these tests replay every operation on
the oracle (CPU) and on libhegpu.so (GPU) with the file's own keys and compare bit for bit.  Without a file the pin
tests SKIP -- parity stays "unpinned" -- while the self-tests below prove on oracle-written vectors of the same
format that the harness reads the container, drives all the operations and notices a single differing word."""
import glob
import os

import numpy as np
import pytest

import hegvec

HERE = os.path.dirname(os.path.abspath(__file__))


def seal_vector_files():
    files = sorted(glob.glob(os.path.join(HERE, "golden", "seal_vectors*.hegvec")))
    env = os.environ.get("HEGPU_SEAL_VECTORS")
    if env and os.path.exists(env):
        files.append(env)
    return files


@pytest.fixture(scope="module")
def selftest_vectors(tmp_path_factory):
    return hegvec.read(hegvec.write_with_oracle(str(tmp_path_factory.mktemp("hegvec") / "selftest.hegvec")))


def test_container_roundtrip(tmp_path):
    rec = {"a": np.arange(24, dtype=np.uint64).reshape(2, 3, 4), "s": np.array([2.0**40]), "steps": np.array([1, -5, 0], dtype=np.int64)}
    p = str(tmp_path / "x.hegvec")
    hegvec.write(p, rec)
    back = hegvec.read(p)
    assert list(back) == list(rec)
    for k in rec:
        assert back[k].dtype == rec[k].dtype and np.array_equal(back[k], rec[k])


def test_harness_selftest_on_oracle_vectors(selftest_vectors):
    vec = selftest_vectors
    names = [n for n, _ in hegvec.operations(vec)]
    assert {"add", "sub", "negate", "multiply", "square", "relinearize", "rescale", "mod_switch", "multiply_plain", "add_plain",
            "sub_plain", "rotate.7", "rotate_low.-5"} <= set(names)
    assert hegvec.run_pin(vec, hegvec.OracleEvaluator(vec)) == []
    bad = dict(vec)
    bad["rotate.7"] = vec["rotate.7"].copy()
    bad["rotate.7"][1, 0, 5] ^= np.uint64(1)  # one bit of one word
    assert hegvec.run_pin(bad, hegvec.OracleEvaluator(bad)) == ["rotate.7"]


@pytest.mark.parametrize("path", seal_vector_files() or [None])
def test_oracle_pinned_by_seal_vectors(path):
    if path is None:
        pytest.skip("no SEAL 4.1 vector file (tools/seal_golden.cpp output) under tests/golden/: parity vs real SEAL stays unpinned")
    vec = hegvec.read(path)
    assert hegvec.run_pin(vec, hegvec.OracleEvaluator(vec)) == []


@pytest.mark.gpu
def test_gpu_harness_selftest_on_oracle_vectors(selftest_vectors):
    import hegpu_loader

    vec = selftest_vectors
    assert hegvec.run_pin(vec, hegvec.GpuEvaluator(vec, hegpu_loader.load())) == []


@pytest.mark.gpu
@pytest.mark.parametrize("path", seal_vector_files() or [None])
def test_gpu_pinned_by_seal_vectors(path):
    if path is None:
        pytest.skip("no SEAL 4.1 vector file (tools/seal_golden.cpp output) under tests/golden/")
    import hegpu_loader

    vec = hegvec.read(path)
    assert hegvec.run_pin(vec, hegvec.GpuEvaluator(vec, hegpu_loader.load())) == []
