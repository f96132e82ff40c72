"""hdf5lite (the reference's HDF5 results files, cosmomodels.py:236-402 and
1552-1677) -- CPU tests.

Pinning: the reader is checked against the one libhdf5-written file that exists in this image
(SciPy's MATLAB v7.3 test file); the writer against that reader plus a structural audit of the
bytes it produced (alignment, fixed node sizes, no overlapping structures, key order, end-of-file
address) following the HDF5 File Format Specification.  No libhdf5 / PyTables is available to
open the written files: that part stays unpinned."""
import os
import struct

import numpy as np
import pytest

from pyflation_b200 import hdf5lite as h

UNDEF = h.UNDEF


def same(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def scipy_hdf5_file():
    import scipy.io
    return os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")


def audit(filename):
    """Walk every structure of a file written by hdf5lite and check what the specification
    fixes: 8-byte alignment, signatures, fixed node sizes, sorted keys, no two structures
    overlapping, everything inside the end-of-file address, which equals the file size."""
    raw = open(filename, "rb").read()
    assert raw[:8] == h.SIGNATURE
    assert raw[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])
    leaf_k, internal_k, flags = struct.unpack_from("<HHI", raw, 16)
    assert (leaf_k, internal_k, flags) == (4, 16, 0)
    base, free, eof, drv = struct.unpack_from("<4Q", raw, 24)
    assert (base, free, drv) == (0, UNDEF, UNDEF) and eof == len(raw)
    extents = [(0, 96, "superblock")]

    def claim(a, n, what):
        assert a % 8 == 0, (what, a)
        assert a + n <= eof, (what, a, n)
        extents.append((a, a + n, what))

    def object_header(addr):
        ver, nmsg, ref, size = struct.unpack_from("<BxHII", raw, addr)
        assert (ver, ref) == (1, 1) and size % 8 == 0
        claim(addr, 16 + size, "object header")
        msgs, p = [], addr + 16
        while p < addr + 16 + size:
            t, n, fl = struct.unpack_from("<HHB", raw, p)
            assert n % 8 == 0 and raw[p + 5:p + 8] == b"\0\0\0"
            msgs.append((t, p + 8, n))
            p += 8 + n
        assert p == addr + 16 + size and len(msgs) == nmsg
        return msgs

    def group(ohdr, btree, heap):
        msgs = object_header(ohdr)
        stab = [m for m in msgs if m[0] == 0x11]
        assert len(stab) == 1 and struct.unpack_from("<QQ", raw, stab[0][1]) == (btree, heap)
        assert raw[heap:heap + 8] == b"HEAP\0\0\0\0"
        seg, freeoff, data = struct.unpack_from("<QQQ", raw, heap + 8)
        claim(heap, 32, "heap header")
        claim(data, seg, "heap data")
        assert seg % 8 == 0 and raw[data:data + 8] == b"\0" * 8
        nxt, fsize = struct.unpack_from("<QQ", raw, data + freeoff)       # the one free block
        assert nxt == 1 and freeoff + fsize == seg and fsize >= 16

        def name_at(off):
            return raw[data + off:raw.index(b"\0", data + off)]

        assert raw[btree:btree + 6] == b"TREE\0\0"
        nent, left, right = struct.unpack_from("<HQQ", raw, btree + 6)
        assert (left, right) == (UNDEF, UNDEF) and nent <= 2 * internal_k
        claim(btree, 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8, "group btree node")
        keys = [struct.unpack_from("<Q", raw, btree + 24 + 16 * i)[0] for i in range(nent + 1)]
        assert name_at(keys[0]) == b""
        names = []
        for i in range(nent):
            snod = struct.unpack_from("<Q", raw, btree + 32 + 16 * i)[0]
            assert raw[snod:snod + 6] == b"SNOD\1\0"
            nsym = struct.unpack_from("<H", raw, snod + 6)[0]
            assert 1 <= nsym <= 2 * leaf_k and (nent == 1 or nsym >= leaf_k)
            claim(snod, 8 + 2 * leaf_k * 40, "symbol table node")
            for j in range(nsym):
                noff, child, cache, _r, s0, s1 = struct.unpack_from("<QQIIQQ", raw, snod + 8 + 40 * j)
                names.append(name_at(noff))
                if cache == 1:
                    group(child, s0, s1)
                else:
                    assert cache == 0
                    dataset(child)
            assert name_at(keys[i + 1]) == names[-1]              # key i+1 = largest name of child i
        assert names == sorted(names) and len(set(names)) == len(names)
        return names

    def dataset(ohdr):
        msgs = dict((t, (p, n)) for t, p, n in object_header(ohdr))
        assert {0x01, 0x03, 0x05, 0x08} <= set(msgs)
        p = msgs[0x01][0]
        assert raw[p] == 1
        rank = raw[p + 1]
        dims = struct.unpack_from("<%dQ" % rank, raw, p + 8)
        dt, used = h.decode_datatype(raw, msgs[0x03][0])
        assert used <= msgs[0x03][1]
        p = msgs[0x08][0]
        assert raw[p] == 3
        if raw[p + 1] == 1:
            addr, size = struct.unpack_from("<QQ", raw, p + 2)
            assert size == int(np.prod(dims)) * dt.itemsize
            if size:
                claim(addr, size, "contiguous data")
            else:
                assert addr == UNDEF
            return
        assert raw[p + 1] == 2 and raw[p + 2] == rank + 1
        root = struct.unpack_from("<Q", raw, p + 3)[0]
        cdims = struct.unpack_from("<%dI" % (rank + 1), raw, p + 11)
        assert cdims[-1] == dt.itemsize
        ksize = 8 + 8 * (rank + 1)
        nsize = 24 + 65 * ksize + 64 * 8
        chunks = []

        levels = {}

        def node(a):
            assert raw[a:a + 5] == b"TREE\1"
            level = raw[a + 5]
            nent, left, right = struct.unpack_from("<HQQ", raw, a + 6)
            assert nent <= 64
            levels.setdefault(level, []).append((a, left, right))
            claim(a, nsize, "chunk btree node")
            keys = []
            for i in range(nent + 1):
                q = a + 24 + i * (ksize + 8)
                keys.append(struct.unpack_from("<II%dQ" % (rank + 1), raw, q))
            assert [k[2:] for k in keys] == sorted(k[2:] for k in keys)
            for i in range(nent):
                child = struct.unpack_from("<Q", raw, a + 24 + i * (ksize + 8) + ksize)[0]
                if level:
                    assert node(child) == keys[i][2:]             # key i = first key of subtree i
                else:
                    nbytes, mask = keys[i][:2]
                    assert mask == 0 and all(o % c == 0 for o, c in zip(keys[i][2:2 + rank], cdims))
                    claim(child, nbytes, "chunk")
                    chunks.append(keys[i][2:2 + rank])
            return keys[0][2:]

        node(root)
        for lvl, nodes in levels.items():                         # siblings: a doubly linked chain
            assert [n[1] for n in nodes] == [UNDEF] + [n[0] for n in nodes[:-1]]
            assert [n[2] for n in nodes] == [n[0] for n in nodes[1:]] + [UNDEF]
        grid = [range(0, d, c) for d, c in zip(dims, cdims)]
        want = sorted(np.array(np.meshgrid(*grid, indexing="ij")).reshape(rank, -1).T.tolist()) \
            if all(len(g) for g in grid) else []
        assert sorted(list(c) for c in chunks) == want            # every chunk of the grid, once

    root_ohdr, cache, _r, bt, hp = struct.unpack_from("<QIIQQ", raw, 64)
    assert cache == 1
    group(root_ohdr, bt, hp)
    extents.sort()
    for (a0, a1, w0), (b0, b1, w1) in zip(extents, extents[1:]):
        assert a1 <= b0, "overlap: %s [%d, %d) and %s [%d, %d)" % (w0, a0, a1, w1, b0, b1)
    return extents


def test_reader_against_the_libhdf5_written_file():
    """User block, superblock v0, symbol table group, v1 object header, layout message v2,
    IEEE double datatype, a string attribute: parsed as SciPy's own test expects
    (scipy/io/matlab/tests: testdouble = pi/4 * arange(9))."""
    fn = scipy_hdf5_file()
    if not os.path.isfile(fn):
        pytest.skip("SciPy's HDF5 test file is not installed")
    assert h.is_hdf5_file(fn)
    with h.open_file(fn) as f:
        assert list(f.root._v_children) == ["testdouble"]
        node = f.root.testdouble
        assert node.shape == (9, 1) and node.dtype == np.float64 and node.kind == "ARRAY"
        assert node.attrs["MATLAB_class"] == "double"
        assert np.array_equal(node[:, 0], np.pi / 4 * np.arange(9))
        assert np.array_equal(node.memmap()[2:5, 0], np.pi / 4 * np.arange(2, 5))
        assert "testdouble" in f.root and "other" not in f.root
        with pytest.raises(h.NoSuchNodeError):
            f.root.other


def test_datatype_messages_round_trip_and_match_libhdf5_bytes():
    # the double of the libhdf5-written file, byte for byte (its dataset header, offset 0x5f8)
    want = bytes.fromhex("11203f0008000000" "000040003 40b0034ff030000".replace(" ", ""))
    assert h.encode_datatype(np.float64) == want
    table = np.dtype([("classname", "S255"), ("datetime", "f8"), ("nfields", "i4"), ("tstart", "f8", (3,))])
    for dt in (np.float64, np.float32, np.int64, np.int32, np.uint8, np.int8, np.complex128, np.complex64,
               np.dtype("S50"), np.bool_, table, np.dtype([("name", "S255"), ("value", "f8")])):
        enc = h.encode_datatype(dt)
        back, used = h.decode_datatype(enc)
        assert used == len(enc) and back == np.dtype(dt), (dt, back)
    # bool: PyTables' 8-bit bit field is what is written; h5py's FALSE / TRUE enum is read as well
    assert h.encode_datatype(np.bool_) == struct.pack("<B3BIHH", 0x14, 0, 0, 0, 1, 0, 8)
    h5py_bool = (struct.pack("<B3BI", 0x18, 2, 0, 0, 1) + h.encode_datatype(np.int8)
                 + b"FALSE\0\0\0" + b"TRUE\0\0\0\0" + b"\0\1")
    assert h.decode_datatype(h5py_bool) == (np.dtype(np.bool_), len(h5py_bool))
    # PyTables' complex: a compound of two doubles called r and i
    enc = h.encode_datatype(np.complex128)
    assert enc[0] == 0x16 and enc[4:8] == struct.pack("<I", 16) and enc[8:9] == b"r" and b"i\0" in enc
    with pytest.raises(NotImplementedError):
        h.encode_datatype(np.dtype(">f8"))
    with pytest.raises(h.HDF5FormatError):
        h.decode_datatype(struct.pack("<B3BI", 0x19, 0, 0, 0, 16) + b"\0" * 16)     # variable length


@pytest.fixture
def sample(tmp_path):
    rng = np.random.default_rng(0)
    y = rng.standard_normal((300, 5, 77)) + 1j * rng.standard_normal((300, 5, 77))
    y[:100, 3:, 40:] = np.nan
    return str(tmp_path / "run.hf5"), y, np.linspace(1e-60, 1e-58, 77)


def write_reference_structure(fn, y, k, filters=None, chunkshape=(2, 5, 77)):
    """The calls createhdf5structure / saveresultsinhdf5 make (cosmomodels.py:277-376)."""
    f = h.openFile(fn, "w")
    g = f.createGroup(f.root, "results", "Results of simulation")
    tres = f.createEArray(g, "tresult", h.Float64Atom(), (0,), filters=filters, expectedrows=8194)
    params = {"solver": h.StringCol(50), "classname": h.StringCol(255), "tstart": h.Float64Col(()),
              "simtstart": h.Float64Col(), "tend": h.Float64Col(), "tstep_wanted": h.Float64Col(),
              "datetime": h.Float64Col(), "nfields": h.IntCol()}
    pt = f.createTable(g, "parameters", params, filters=filters)
    pp = f.createTable(g, "pot_params", {"name": h.StringCol(255), "value": h.Float64Col()}, filters=filters)
    bg = f.createGroup(f.root, "bgresults", "Background results")
    f.createArray(bg, "tresult", np.arange(50) * 0.01)
    f.createArray(bg, "yresult", np.arange(150.0).reshape(50, 3, 1))
    ye = f.createEArray(g, "yresult", h.ComplexAtom(itemsize=16), (0,) + y.shape[1:], filters=filters,
                        expectedrows=8194, chunkshape=chunkshape)
    f.createArray(g, "k", k)
    f.createArray(g, "fotstartindex", np.arange(len(k)))
    f.createArray(g, "empty", np.zeros((0, 3)))
    row = pt.row
    for key, v in dict(solver="rkdriver_tsix", classname="FOCanonicalTwoStage", tstart=0.0, simtstart=0.0,
                       tend=83.0, tstep_wanted=0.01, nfields=1, datetime=20260101.0).items():
        row[key] = v
    row.append()
    pt.flush()
    for name, val in (("mass", 6.3e-6), ("nfields", 1)):
        r = pp.row
        r["name"], r["value"] = name, val
        r.append()
    ye.append(y[:123])
    ye.append(y[123:])
    tres.append(np.arange(len(y)) * 0.01)
    f.flush()
    f.close()


@pytest.mark.parametrize("filters", [None, h.Filters(complevel=2, complib="zlib"),
                                     h.Filters(complevel=1, complib="zlib", shuffle=False)])
def test_reference_file_structure_round_trip(sample, filters):
    fn, y, k = sample
    write_reference_structure(fn, y, k, filters)
    audit(fn)
    with h.openFile(fn, "r") as f:
        res = f.root.results
        assert sorted(res._v_children) == ["empty", "fotstartindex", "k", "parameters", "pot_params",
                                           "tresult", "yresult"]
        assert res._v_attrs["TITLE"] == "Results of simulation" and res._v_attrs["CLASS"] == "GROUP"
        assert f.root._v_attrs["PYTABLES_FORMAT_VERSION"] == "2.1"
        assert res.yresult.kind == "EARRAY" and res.yresult.attrs["EXTDIM"] == 0
        assert res.yresult.maxshape == (None, 5, 77) and res.yresult.chunkshape == (2, 5, 77)
        assert same(res.yresult[:], y) and same(res.yresult.read(), y)
        assert same(res.yresult[7:250:3, 1, ::2], y[7:250:3, 1, ::2])
        assert same(res.yresult[-1], y[-1]) and same(res.yresult[::-2, 2], y[::-2, 2])
        assert same(res.yresult[..., 5], y[..., 5]) and same(res.yresult[299:300], y[299:])
        assert same(res.yresult[[3, 1, 200]], y[[3, 1, 200]])
        with pytest.raises(IndexError):
            res.yresult[300]
        assert same(res.tresult[:], np.arange(300) * 0.01)
        assert same(res.k[:], k) and same(res.k.memmap(), k) and res.yresult.memmap() is None
        assert res.empty.shape == (0, 3) and res.empty[:].shape == (0, 3)
        assert same(f.root.bgresults.yresult[0], np.arange(3.0).reshape(3, 1))
        par = res.parameters
        assert par.kind == "TABLE" and par.attrs["NROWS"] == 1 and len(par) == 1
        assert par.colnames == sorted(par.colnames) and par.attrs["FIELD_0_NAME"] == "classname"
        assert par[0]["classname"] == b"FOCanonicalTwoStage" and par[0]["nfields"] == 1
        assert par[0]["tend"] == 83.0 and par[0]["solver"] == b"rkdriver_tsix"
        assert [(r["name"], r["value"]) for r in res.pot_params] == [(b"mass", 6.3e-6), (b"nfields", 1.0)]
        assert f.getNode(f.root, "results") is res and f.get_node("/results/k") is res.k
        assert "/results/yresult" in f and "/results/nothing" not in f
    if filters is not None and filters.complevel:
        # NaN prefixes and the regular time axis compress: the file is smaller than the raw data
        assert os.path.getsize(fn) < y.nbytes


def test_append_mode_extends_arrays_and_adds_nodes(sample):
    """saveallresults opens an existing file in "a" mode and appends (cosmomodels.py:247-249,
    270-272); combine_source_and_fofile adds the source term to a copy (sohelpers.py:73-80)."""
    fn, y, k = sample
    write_reference_structure(fn, y, k, h.Filters(complevel=1, complib="zlib"))
    with h.openFile(fn, "a") as f:
        f.root.results.yresult.append(y[:10])
        f.root.results.tresult.append(np.arange(10) * 1.0)
        f.createArray(f.root.results, "sourceterm", y[:, 0, :])
        f.root.results.pot_params.append([("extra", 2.5)])
        with pytest.raises(h.NodeError):
            f.createArray(f.root.results, "k", k)
    audit(fn)
    with h.openFile(fn) as f:
        res = f.root.results
        assert res.yresult.shape == (310, 5, 77) and res.tresult.shape == (310,)
        assert same(res.yresult[:300], y) and same(res.yresult[300:], y[:10]) and same(res.yresult[295:305],
                                                                                         np.concatenate([y[295:], y[:5]]))
        assert same(res.sourceterm[:], y[:, 0, :]) and len(res.pot_params) == 3
        assert res.pot_params[2]["name"] == b"extra" and res.parameters[0]["tend"] == 83.0
        assert res.yresult.filters is not None
        with pytest.raises(IOError):
            f.createGroup(f.root, "nope")
    # an untouched "a" session leaves the file's bytes alone
    before = open(fn, "rb").read()
    h.openFile(fn, "a").close()
    assert open(fn, "rb").read() == before


def test_many_children_many_chunks_and_split_rows(tmp_path):
    """More than one symbol-table node per group (> 8 children), a chunk B-tree of two levels
    (> 64 chunks) and chunks that split the last axis."""
    fn = str(tmp_path / "big.h5")
    rng = np.random.default_rng(1)
    a = rng.standard_normal((200, 10))
    with h.open_file(fn, "w", title="many") as f:
        for i in range(41):
            f.create_array("/", "node%03d" % i, np.full(i, i, dtype=np.int32), title="n%d" % i)
        g = f.create_group("/", "sub")
        f.create_group(g, "subsub", "deep")
        f.create_array("/sub/subsub", "scalar", np.float64(2.5))
        f.create_array("/sub/subsub", "flags", np.array([True, False, True]))
        e = f.create_earray("/sub", "split", h.Float64Atom(), (0, 10), chunkshape=(1, 4))
        e.append(a)
        f.create_earray("/sub", "never_filled", h.Int64Atom(), (0, 7))
    ext = audit(fn)
    assert sum(1 for e in ext if e[2] == "chunk btree node") >= 11      # 600 chunks: 10 leaves + a root
    with h.open_file(fn) as f:
        assert f.root._v_attrs["TITLE"] == "many" and len(f.root) == 42
        for i in range(41):
            node = getattr(f.root, "node%03d" % i)
            assert same(node[:], np.full(i, i, dtype=np.int32)) and node.title == "n%d" % i
        assert f.root.sub.subsub.scalar.read() == 2.5 and f.root.sub.subsub.scalar.shape == ()
        assert f.root.sub.subsub.flags[:].tolist() == [True, False, True]
        assert same(f.root.sub.split[:], a) and same(f.root.sub.split[17:150, 3:9], a[17:150, 3:9])
        assert f.root.sub.never_filled.shape == (0, 7) and f.root.sub.never_filled[:].shape == (0, 7)
        assert [n._v_name for n in f.walk_nodes("/sub")] == ["sub", "never_filled", "split", "subsub", "flags", "scalar"]
    assert h.plan_chunkshape((0, 5, 2053), 16) == (25, 5, 2053)
    assert h.plan_chunkshape((0, 145, 1 << 20), 16) == (1, 145, 349526)          # 2.4 GB rows: three pieces along k
    assert h.plan_chunkshape((0,), 8) == (1 << 19,)


def test_errors(tmp_path):
    bad = tmp_path / "bad.h5"
    bad.write_bytes(b"not an hdf5 file" * 20)
    assert not h.is_hdf5_file(str(bad))
    with pytest.raises(h.HDF5FormatError):
        h.open_file(str(bad))
    with pytest.raises(IOError):
        h.open_file(str(tmp_path / "missing" / "x.h5"), "w")
    with pytest.raises(IOError):
        h.open_file(str(tmp_path / "missing.h5"), "r")
    with pytest.raises(NotImplementedError):
        h.Filters(complevel=2, complib="blosc")               # not an HDF5 codec: refuse, do not guess
    assert h.Filters(complevel=0, complib="blosc").complevel == 0
    # a dataset stored with a foreign filter is reported by name
    fn = str(tmp_path / "f.h5")
    with h.open_file(fn, "w") as f:
        f.create_earray("/", "x", h.Float64Atom(), (0,), filters=h.Filters(complevel=1)).append(np.arange(9.0))
    raw = bytearray(open(fn, "rb").read())
    at = raw.index(b"deflate\0") - 8
    raw[at:at + 2] = struct.pack("<H", 32001)
    open(fn, "wb").write(bytes(raw))
    with h.open_file(fn) as f:
        with pytest.raises(h.HDF5FormatError, match="blosc"):
            f.root.x[:]
    # version-2 superblocks are named, not misparsed
    raw[8] = 2
    open(fn, "wb").write(bytes(raw))
    with pytest.raises(h.HDF5FormatError, match="superblock version 2"):
        h.open_file(fn)
    f = h.open_file(str(tmp_path / "c.h5"), "w")
    f.close()
    with pytest.raises(h.ClosedFileError):
        f.create_group("/", "g")


def test_reader_follows_object_header_continuation_blocks(tmp_path):
    """libhdf5 moves messages that no longer fit (attributes added later: PyTables' TITLE, CLASS,
    ...) into continuation blocks (message 0x0010).  The writer never needs one, so a file of its
    own is rearranged by hand here: the last attribute of a dataset's header is moved to the end
    of the file, a continuation message plus a NIL message take its place."""
    fn = str(tmp_path / "c.h5")
    with h.open_file(fn, "w") as f:
        node = f.create_array("/", "x", np.arange(5.0), title="moved title")
        node.attrs["zz_last"] = np.arange(3, dtype=np.int32)
    raw = bytearray(open(fn, "rb").read())
    with h.open_file(fn) as f:
        addr = [a for n, a, c in f._img.group_entries(*struct.unpack_from("<QQ", raw, 80)) if n == "x"][0]
        before = dict(f.root.x.attrs)
    assert before["FLAVOR"] == "numpy" and before["zz_last"].tolist() == [0, 1, 2]
    nmsg, _ref, size = struct.unpack_from("<HII", raw, addr + 2)
    p, msgs = addr + 16, []
    while p < addr + 16 + size:
        t, n = struct.unpack_from("<HH", raw, p)
        msgs.append((t, p, 8 + n))
        p += 8 + n
    t, at, length = msgs[-1]
    assert t == 0x0C and length >= 32                          # an attribute: the last message
    moved = bytes(raw[at:at + length])
    new_at = len(raw)
    raw[at:at + length] = (struct.pack("<HHB3x", 0x10, 16, 0) + struct.pack("<QQ", new_at, length)
                           + struct.pack("<HHB3x", 0, length - 32, 0) + b"\0" * (length - 32))
    raw += moved
    struct.pack_into("<H", raw, addr + 2, nmsg + 2)
    struct.pack_into("<Q", raw, 40, len(raw))                  # end-of-file address
    open(fn, "wb").write(bytes(raw))
    with h.open_file(fn) as f:
        after = dict(f.root.x.attrs)
        assert sorted(after) == sorted(before) and f.root.x.title == "moved title"
        assert all(np.array_equal(after[q], before[q]) for q in before)
        assert np.array_equal(f.root.x[:], np.arange(5.0))


def test_nodes_of_a_closed_file(tmp_path):
    fn = str(tmp_path / "closed.h5")
    with h.open_file(fn, "w") as f:
        f.create_earray("/", "chunked", h.Float64Atom(), (0,)).append(np.arange(9.0))
        f.create_array("/", "plain", np.arange(4.0))
    f = h.open_file(fn)
    chunked, plain = f.root.chunked, f.root.plain
    f.close()
    with pytest.raises(h.ClosedFileError):
        chunked[:]
    assert plain[:].tolist() == [0.0, 1.0, 2.0, 3.0]         # contiguous data: its own memory map
    with pytest.raises(h.ClosedFileError):
        f.get_node("/plain")


def test_random_trees_round_trip_and_pass_the_audit(tmp_path):
    """Property test (hypothesis): random group trees of arrays / extendable arrays / tables of
    random dtypes, shapes, chunk shapes and filters are read back bit for bit and their bytes
    satisfy the structural audit."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    dtypes = st.sampled_from([np.float64, np.float32, np.int64, np.int32, np.uint8, np.complex128,
                              np.complex64, np.dtype("S7"), np.bool_])
    names = st.text(alphabet="abcdefghijklmnopqrstuvwxyz_0123456789", min_size=1, max_size=12)

    def array_of(dt, shape, seed):
        rng = np.random.default_rng(seed)
        dt = np.dtype(dt)
        n = int(np.prod(shape))
        if dt.kind == "c":
            a = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        elif dt.kind == "S":
            a = np.array([b"x" * int(v) for v in rng.integers(0, dt.itemsize + 1, n)], dtype=dt)
        elif dt.kind == "b":
            a = rng.integers(0, 2, n).astype(bool)
        elif dt.kind in "iu":
            a = rng.integers(0, 100, n)
        else:
            a = rng.standard_normal(n)
        return np.asarray(a).astype(dt).reshape(shape)

    leaf = st.tuples(st.sampled_from(["array", "earray", "table"]), dtypes,
                     st.lists(st.integers(0, 5), min_size=0, max_size=3),
                     st.lists(st.integers(1, 4), min_size=3, max_size=3),
                     st.sampled_from([None, (1, True), (6, False)]), st.integers(0, 2 ** 31))
    tree = st.dictionaries(names, st.one_of(leaf, st.dictionaries(names, leaf, max_size=11)), max_size=12)
    counter = [0]

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(tree)
    def check(spec):
        counter[0] += 1
        fn = str(tmp_path / ("t%d.h5" % counter[0]))
        want = {}

        def add(f, where, path, name, item):
            kind, dt, shape, chunk, filt, seed = item
            filters = None if filt is None else h.Filters(complevel=filt[0], shuffle=filt[1])
            if kind == "array":
                data = array_of(dt, tuple(shape), seed)
                f.create_array(where, name, data)
            elif kind == "earray":
                shape = tuple(shape) or (3,)
                data = array_of(dt, shape, seed)
                node = f.create_earray(where, name, h.Atom(dt), (0,) + shape[1:], filters=filters,
                                       chunkshape=tuple(chunk[:len(shape)]))
                cut = shape[0] // 2
                node.append(data[:cut])
                node.append(data[cut:])
            else:
                rows = (shape or [2])[0]
                tdt = np.dtype([("a_" + np.dtype(dt).name, dt), ("value", "f8"), ("vec", "i4", (2,))])
                data = np.zeros(rows, dtype=tdt)
                data[tdt.names[0]] = array_of(dt, (rows,), seed)
                data["value"] = np.arange(rows) * 0.5
                data["vec"] = np.arange(2 * rows).reshape(rows, 2)
                node = f.create_table(where, name, tdt, filters=filters, chunkshape=(chunk[0],))
                node.append(data)
            want[path + "/" + name] = data

        with h.open_file(fn, "w") as f:
            for name, item in spec.items():
                if isinstance(item, dict):
                    g = f.create_group("/", name, "group " + name)
                    for sub, subitem in item.items():
                        add(f, g, "/" + name, sub, subitem)
                else:
                    add(f, "/", "", name, item)
        audit(fn)
        with h.open_file(fn) as f:
            found = [p for p in want if p in f]
            assert sorted(found) == sorted(want)
            for path, data in want.items():
                got = f.get_node(path).read()
                assert same(got, data), (path, got.dtype, data.dtype)
        os.remove(fn)

    check()
    assert counter[0] >= 20                                   # the property really ran


def test_decoders_of_message_versions_the_writer_never_emits():
    """Versions libhdf5 >= 1.8 may choose (dataspace v2, attribute v2 / v3, compound v3, array v3,
    filter pipeline v2), hand-encoded from the specification."""
    from pyflation_b200.hdf5lite import _decode_dataspace, _decode_attribute, _decode_filters
    f8 = h.encode_datatype(np.float64)
    # dataspace v2: version, rank, flags (max dims present), type (1 = simple), dims, max dims
    ds2 = struct.pack("<BBBB", 2, 2, 1, 1) + struct.pack("<4Q", 3, 5, h.UNDEF, 5)
    assert _decode_dataspace(ds2, 0) == ((3, 5), (None, 5))
    assert _decode_dataspace(struct.pack("<BBBB", 2, 0, 0, 2), 0) == (None, None)        # null dataspace
    assert _decode_dataspace(struct.pack("<BBBB", 2, 0, 0, 0), 0) == ((), ())            # scalar
    # attribute v3: version, flags, name / datatype / dataspace sizes, character set; nothing padded
    name, space = b"answer\0", struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 2)
    a3 = struct.pack("<BBHHHB", 3, 0, len(name), len(f8), len(space), 0) + name + f8 + space + \
        np.array([1.5, -2.0]).tobytes()
    got = _decode_attribute(a3, 0)
    assert got[0] == "answer" and got[1].tolist() == [1.5, -2.0]
    a2 = struct.pack("<BBHHH", 2, 0, len(name), len(f8), len(space)) + name + f8 + space + \
        np.array([1.5, -2.0]).tobytes()
    assert _decode_attribute(a2, 0)[1].tolist() == [1.5, -2.0]
    # compound v3: names not padded, member offsets in the fewest bytes that hold the size (1 here)
    i4 = h.encode_datatype(np.int32)
    arr3 = struct.pack("<B3BI", 0x3A, 0, 0, 0, 24) + struct.pack("<B", 1) + struct.pack("<I", 3) + f8
    c3 = (struct.pack("<B3BI", 0x36, 2, 0, 0, 28) + b"n\0" + b"\0" + i4 + b"vec\0" + b"\x04" + arr3)
    dt, used = h.decode_datatype(c3)
    assert used == len(c3) and dt == np.dtype([("n", "<i4"), ("vec", "<f8", (3,))])
    # filter pipeline v2: no name for the predefined filters, nothing padded
    p2 = struct.pack("<BB", 2, 2) + struct.pack("<HHH", 2, 1, 1) + struct.pack("<I", 8) + \
        struct.pack("<HHH", 1, 1, 1) + struct.pack("<I", 4)
    assert _decode_filters(p2, 0) == [(2, 1, (8,)), (1, 1, (4,))]
