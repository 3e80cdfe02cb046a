"""Data-file bridge: the reference's save_to_file / load_from_file format
(include/xerus/misc/fileIO.h:102-164, tensor.cpp:1781-1845, tensorNetwork.cpp:1429-1505, ttNetwork.cpp:1455-1488).

tests/golden/files/* were written by the unmodified reference (oracle/drivers/ref_files.cpp, `ref_files write`);
contents.bin holds the same objects through the record container of the other goldens.  The CPU tests use the host-side
xb_file_* entry points only (no device); where oracle/_ref is built (this container) files written by libxb200 are also
loaded back by the reference's own reader.  The GPU test moves a file into a device-resident TT and back."""
import os
import subprocess
import sys

import numpy as np
import pytest

import xerus_b200 as xb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = os.path.join(ROOT, "tests", "golden", "files")
REF_FILES = os.path.join(ROOT, "oracle", "_ref", "ref_files")
sys.path.insert(0, os.path.join(ROOT, "oracle"))


@pytest.fixture(scope="module")
def contents():
    from make_golden import read_container
    return read_container(os.path.join(FILES, "contents.bin"))


def tt_record(rec, name):
    d = int(rec[name + ".d"])
    core = int(rec[name + ".core"])
    return [rec["%s.c%d" % (name, i)] for i in range(d)], core


@pytest.mark.parametrize("fname,name,kind", [("tttensor.bin", "tttensor", "TTTensor"), ("tttensor_core2.bin", "tttensor_core2", "TTTensor"),
                                             ("tttensor_sum.bin", "tttensor_sum", "TTTensor"), ("ttoperator.bin", "ttoperator", "TTOperator"),
                                             ("laplace.bin", "laplace", "TTOperator")])
def test_read_reference_binary_tt(contents, fname, name, kind):
    f = xb.read_file(os.path.join(FILES, fname))
    cores, core = tt_record(contents, name)
    assert f["kind"] == kind
    assert f["canonicalized"] == (core >= 0)
    if core >= 0:
        assert f["core_position"] == core
    assert f["ranks"] == [c.shape[-1] for c in cores[:-1]]
    assert len(f["components"]) == len(cores)
    for a, b in zip(f["components"], cores):
        assert a.shape == b.shape and np.array_equal(a, b)          # binary storage is exact


@pytest.mark.parametrize("fname,name", [("tttensor.tsv", "tttensor"), ("ttoperator.tsv", "ttoperator")])
def test_read_reference_tsv_tt(contents, fname, name):
    f = xb.read_file(os.path.join(FILES, fname))
    cores, core = tt_record(contents, name)
    assert f["canonicalized"] == (core >= 0)
    for a, b in zip(f["components"], cores):
        # the reference prints digits10 + 1 = 16 significant digits (tensor.cpp:1783): one short of a bit-exact round trip
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-15, atol=0)


def test_read_reference_tensors(contents):
    for fname, name, tol in [("tensor_dense.bin", "tensor_dense", 0), ("tensor_sparse.bin", "tensor_sparse", 0),
                             ("tensor_dense.tsv", "tensor_dense", 1e-15), ("tensor_sparse.tsv", "tensor_sparse", 1e-15)]:
        f = xb.read_file(os.path.join(FILES, fname))
        assert f["kind"] == "Tensor" and list(f["data"].shape) == list(contents[name].shape)
        if tol == 0:
            assert np.array_equal(f["data"], contents[name])          # the lazy factor is applied in the file (tensor.cpp:1795,:1802)
        else:
            assert np.allclose(f["data"], contents[name], rtol=tol, atol=0)


@pytest.mark.parametrize("fname", ["tttensor.bin", "tttensor_core2.bin", "tttensor_sum.bin", "ttoperator.bin", "laplace.bin",
                                   "tttensor.tsv", "ttoperator.tsv", "tensor_dense.bin", "tensor_dense.tsv"])
def test_writer_reproduces_reference_files_byte_for_byte(tmp_path, fname):
    src = os.path.join(FILES, fname)
    f = xb.read_file(src)
    fmt = xb.FileFormat.TSV if fname.endswith(".tsv") else xb.FileFormat.BINARY
    out = str(tmp_path / fname)
    if f["kind"] == "Tensor":
        xb.write_tensor_file(out, f["data"], fmt)
    else:
        xb.write_tt_file(out, f["components"], f["canonicalized"], f["core_position"], fmt)
    if fmt == xb.FileFormat.BINARY:
        assert open(out, "rb").read() == open(src, "rb").read()
    else:
        # text: same tokens in the same order (the values were parsed from 16 digits and are printed with 16 digits again)
        assert open(out).read().split() == open(src).read().split()


def test_rejects_what_the_reference_rejects(tmp_path):
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"Xerus xerus::Foo datafile.\nFormat: Binary\n")
    with pytest.raises(xb.XerusError):
        xb.read_file(str(bad))
    raw = open(os.path.join(FILES, "tttensor.bin"), "rb").read()
    (tmp_path / "short.bin").write_bytes(raw[:-9])                    # truncated: "Unexpected end of stream"
    with pytest.raises(xb.XerusError):
        xb.read_file(str(tmp_path / "short.bin"))
    hdr = raw.index(b"Binary\n") + 7
    (tmp_path / "ver.bin").write_bytes(raw[:hdr] + (2).to_bytes(8, "little") + raw[hdr + 8:])   # version 2 (ttNetwork.cpp:1477)
    with pytest.raises(xb.XerusError):
        xb.read_file(str(tmp_path / "ver.bin"))
    with pytest.raises(xb.XerusError):
        xb.read_file(str(tmp_path / "does_not_exist.bin"))
    with pytest.raises(xb.XerusError):                                # neighbouring bonds must agree
        xb.write_tt_file(str(tmp_path / "x.bin"), [np.zeros((1, 2, 3)), np.zeros((2, 2, 1))])


@pytest.mark.skipif(not os.path.exists(REF_FILES), reason="oracle/_ref not built (needs the reference tree)")
@pytest.mark.parametrize("fmt", [0, 1])
def test_reference_reads_files_written_here(tmp_path, fmt):
    rng = np.random.default_rng(11)
    cores = [rng.standard_normal(s) for s in [(1, 3, 4), (4, 2, 5), (5, 4, 2), (2, 3, 1)]]
    out = str(tmp_path / "x.dat")
    xb.write_tt_file(out, cores, canonicalized=False, core_position=0, fmt=fmt)
    txt = subprocess.run([REF_FILES, "read", out], check=True, capture_output=True, text=True).stdout
    info = dict(line.split(" ", 1) for line in txt.strip().splitlines())
    assert info["kind"] == "TTTensor" and info["dims"].split() == ["3", "2", "4", "3"] and info["ranks"].split() == ["4", "5", "2"]
    full = cores[0]
    for c in cores[1:]:
        full = np.tensordot(full, c, axes=([full.ndim - 1], [0]))
    assert abs(float(info["norm"]) - np.linalg.norm(full)) < 1e-12 * np.linalg.norm(full)
    # operator, canonicalised flag, and a second trip through the reference's writer
    ops = [rng.standard_normal(s) for s in [(1, 2, 3, 3), (3, 2, 2, 1)]]
    out2, out3 = str(tmp_path / "A.dat"), str(tmp_path / "A_ref.dat")
    xb.write_tt_file(out2, ops, canonicalized=True, core_position=1, fmt=fmt)
    subprocess.run([REF_FILES, "copy", out2, out3], check=True)
    back = xb.read_file(out3)
    assert back["kind"] == "TTOperator" and back["canonicalized"] and back["core_position"] == 1
    for a, b in zip(back["components"], ops):
        assert np.allclose(a, b, rtol=1e-15 if fmt else 0, atol=0)
    A = rng.standard_normal((3, 1, 4))
    xb.write_tensor_file(str(tmp_path / "T.dat"), A, fmt)
    txt = subprocess.run([REF_FILES, "read", str(tmp_path / "T.dat")], check=True, capture_output=True, text=True).stdout
    assert "dims 3 1 4" in txt and abs(float(txt.split("norm")[1]) - np.linalg.norm(A)) < 1e-13


@pytest.mark.gpu
def test_load_round_save_on_device(tmp_path, contents):
    from oracle import tt_oracle as O
    x = xb.load_from_file(os.path.join(FILES, "tttensor_core2.bin"))
    assert isinstance(x, xb.TTTensor) and x.ranks() == [3, 6, 4] and x.corePosition == 2 and x.canonicalized
    cores, core = tt_record(contents, "tttensor_core2")
    for i, c in enumerate(cores):
        assert np.array_equal(x.get_component(i), c)
    out = str(tmp_path / "same.bin")
    xb.save_to_file(x, out)
    assert open(out, "rb").read() == open(os.path.join(FILES, "tttensor_core2.bin"), "rb").read()
    ref = O.TT(cores, core_position=core)
    x.round(3)
    ref.round(3)
    xb.save_to_file(x, str(tmp_path / "rounded.tsv"), xb.FileFormat.TSV)
    y = xb.load_from_file(str(tmp_path / "rounded.tsv"))
    assert y.ranks() == ref.ranks() and y.corePosition == x.corePosition
    assert O.tt_distance_rel(O.TT(y.cores(), core_position=y.corePosition), ref) < 1e-9
    A = xb.load_from_file(os.path.join(FILES, "laplace.bin"))
    assert isinstance(A, xb.TTOperator) and A.ranks() == [2, 2, 2]
    T = xb.load_from_file(os.path.join(FILES, "tensor_dense.bin"))
    assert np.array_equal(T, contents["tensor_dense"])
