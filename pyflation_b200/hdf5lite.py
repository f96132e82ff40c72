"""hdf5lite - the part of the HDF5 file format that pyflation's result files use, in NumPy.

The reference stores its results with PyTables (cosmomodels.py:236-402: groups ``results`` /
``bgresults``, ``Array`` / ``EArray`` datasets of float64, int64 and complex128, two small
``Table``s of parameters) and reads them back in ``make_wrapper_model`` (cosmomodels.py:1552-1677)
and ``sohelpers.combine_source_and_fofile`` (sohelpers.py:20-86).  Neither PyTables, h5py nor
libhdf5 exists in this image, so this module implements the on-disk format itself, from the
published HDF5 File Format Specification (version 1.1 / "1.8-compatible" structures):

    superblock version 0/1, version-1 object headers (+ continuation blocks), "old style" groups
    (symbol-table message -> version-1 B-tree -> SNOD nodes -> local heap), dataspace messages
    v1/v2, datatype classes fixed-point / float / string / bitfield / compound (v1-v3) / enum /
    array, data layout messages v1-v3 (compact, contiguous, chunked with a version-1 chunk
    B-tree), the filter pipeline (shuffle, deflate, fletcher32 checksums are skipped), attribute
    messages v1-v3, fill-value and modification-time messages.

This is what PyTables (any version, default settings) and h5py (``libver='earliest'``, the
default) write.  Not implemented, with an explicit error: superblock v2/v3 with "OHDR" object
headers and link messages (HDF5 1.10 ``libver='latest'``), variable-length types, external
storage, and the blosc / lzo / bzip2 filters (files the reference wrote with its default
``complib="blosc"`` need ``ptrepack --complib zlib`` first: the blosc codecs are not part of HDF5).

Pinning (see tests/test_hdf5lite.py and DESIGN.md section 5): the READER is checked against the
one libhdf5-written file that exists in this image (a MATLAB v7.3 file among SciPy's test
data: user block, superblock v0, symbol table, v1 object header, layout v2, attribute); the
WRITER is checked by reading its files back through that reader and by structural checks
(alignment, node sizes, key ordering) against the specification.  No libhdf5 has opened a file
written here: compatibility of the writer with PyTables is by construction, not by test.

The API is the slice of PyTables the reference calls: ``open_file`` (``openFile``),
``create_group / create_array / create_earray / create_table`` (and the camelCase spellings of
PyTables 2.x the reference uses), ``get_node``, natural naming (``f.root.results.yresult``),
``in``, slicing, ``EArray.append``, ``Table.row`` / ``append`` / ``colnames`` / iteration,
``attrs``, ``NoSuchNodeError``, ``Filters``, the ``*Atom`` / ``*Col`` descriptors.
"""
import mmap
import os
import struct
import time
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
_GROUP_LEAF_K = 4            # symbol-table node holds up to 2K entries        (spec III.B)
_GROUP_INTERNAL_K = 16       # group B-tree node holds up to 2K children        (spec III.A.1)
_CHUNK_K = 32                # chunk B-tree node: 2K children (superblock v0 has no field: default)
_CHUNK_TARGET_BYTES = 4 << 20
_CHUNK_MAX_BYTES = 1 << 30   # a chunk's size is a 32-bit field

FILTER_DEFLATE, FILTER_SHUFFLE, FILTER_FLETCHER32 = 1, 2, 3
_FILTER_NAMES = {4: "szip", 5: "nbit", 6: "scaleoffset", 305: "lzo", 307: "bzip2", 32001: "blosc",
                 32026: "blosc2", 32004: "lz4", 32015: "zstd"}


class HDF5FormatError(IOError):
    """The file is not HDF5, is damaged, or uses a part of the format this module leaves out."""


class NodeError(AttributeError, LookupError):
    pass


class NoSuchNodeError(NodeError):
    """tables.NoSuchNodeError"""


class ClosedFileError(ValueError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------------
# descriptors (tables.Atom / tables.Col / tables.Filters)
# ------------------------------------------------------------------------------------------------
class Atom(object):
    def __init__(self, dtype, shape=()):
        base = np.dtype(dtype)
        self.shape = tuple(int(v) for v in np.atleast_1d(shape))
        self.dtype = np.dtype((base, self.shape)) if self.shape else base
        self.base = base

    @classmethod
    def from_dtype(cls, dtype):
        return cls(dtype)


def Float64Atom(shape=()):
    return Atom(np.float64, shape)


def Float32Atom(shape=()):
    return Atom(np.float32, shape)


def Int64Atom(shape=()):
    return Atom(np.int64, shape)


def Int32Atom(shape=()):
    return Atom(np.int32, shape)


def ComplexAtom(itemsize=16, shape=()):
    if itemsize not in (8, 16):
        raise ValueError("ComplexAtom: itemsize must be 8 or 16")
    return Atom(np.complex64 if itemsize == 8 else np.complex128, shape)


def StringAtom(itemsize, shape=()):
    return Atom("S%d" % int(itemsize), shape)


Float64Col, Float32Col, Int64Col, Int32Col, IntCol = Float64Atom, Float32Atom, Int64Atom, Int32Atom, Int32Atom
FloatCol, ComplexCol, StringCol = Float64Atom, ComplexAtom, StringAtom


class Filters(object):
    """tables.Filters: ``complevel`` 0 stores raw chunks; ``complib`` "zlib" is the one codec here."""

    def __init__(self, complevel=0, complib="zlib", shuffle=True, fletcher32=False):
        self.complevel, self.complib, self.shuffle = int(complevel), complib, bool(shuffle)
        if fletcher32:
            raise NotImplementedError("fletcher32 checksums are not written")
        if self.complevel and complib != "zlib":
            raise NotImplementedError("complib=%r: only 'zlib' can be written without libhdf5 plugins "
                                      "(use complevel=0 or complib='zlib')" % (complib,))


# ------------------------------------------------------------------------------------------------
# datatype messages <-> numpy dtypes                                             (spec IV.A.2.d)
# ------------------------------------------------------------------------------------------------
def _cstring(buf, p):
    e = buf.find(b"\0", p)
    if e < 0:
        raise HDF5FormatError("unterminated name in a datatype / heap")
    return bytes(buf[p:e]), e - p


def decode_datatype(buf, p=0):
    """(numpy dtype, bytes consumed) of the datatype message at buf[p:]."""
    b0 = buf[p]
    cls, ver = b0 & 0x0F, b0 >> 4
    bits = buf[p + 1] | (buf[p + 2] << 8) | (buf[p + 3] << 16)
    size = struct.unpack_from("<I", buf, p + 4)[0]
    q = p + 8
    order = ">" if bits & 1 else "<"
    if cls == 0:                                          # fixed point
        dt = np.dtype("%s%s%d" % (order, "i" if bits & 8 else "u", size))
        q += 4
    elif cls == 1:                                        # IEEE float
        dt = np.dtype("%sf%d" % (order, size))
        q += 12
    elif cls == 3:                                        # fixed-length string
        dt = np.dtype("S%d" % size)
    elif cls == 4:                                        # bit field; one byte = PyTables' bool (H5T_STD_B8)
        dt = np.dtype(np.bool_) if size == 1 else np.dtype("%su%d" % (order, size))
        q += 4
    elif cls == 5:                                        # opaque
        q += _pad8(bits & 0xFF)
        dt = np.dtype("V%d" % size)
    elif cls == 7:                                        # object / region reference
        dt = np.dtype("V%d" % size)
    elif cls == 6:                                        # compound
        names, formats, offsets = [], [], []
        for _ in range(bits & 0xFFFF):
            name, n = _cstring(buf, q)
            q += n + 1 if ver >= 3 else ((n + 8) // 8) * 8
            if ver >= 3:
                nb = 1
                while nb < 8 and size >= (1 << (8 * nb)):
                    nb += 1
                off = int.from_bytes(bytes(buf[q:q + nb]), "little")
                q += nb
            else:
                off = struct.unpack_from("<I", buf, q)[0]
                q += 4
            dims = ()
            if ver == 1:
                nd = buf[q]
                dims = struct.unpack_from("<4I", buf, q + 12)[:nd]
                q += 28
            mdt, used = decode_datatype(buf, q)
            q += used
            if dims:
                mdt = np.dtype((mdt, tuple(int(d) for d in dims)))
            names.append(name.decode("utf-8", "replace"))
            formats.append(mdt)
            offsets.append(off)
        if (names == ["r", "i"] and formats[0] == formats[1] and formats[0].kind == "f"
                and offsets == [0, formats[0].itemsize] and size == 2 * formats[0].itemsize):
            dt = np.dtype("%sc%d" % (formats[0].byteorder.replace("=", "<").replace("|", "<"), size))
        else:
            dt = np.dtype(dict(names=names, formats=formats, offsets=offsets, itemsize=size))
    elif cls == 8:                                        # enumeration (h5py's bool: FALSE / TRUE on int8)
        base, used = decode_datatype(buf, q)
        q += used
        nmemb = bits & 0xFFFF
        names = []
        for _ in range(nmemb):
            name, n = _cstring(buf, q)
            q += n + 1 if ver >= 3 else ((n + 8) // 8) * 8
            names.append(name)
        q += nmemb * base.itemsize
        dt = np.dtype(np.bool_) if (base.itemsize == 1 and names == [b"FALSE", b"TRUE"]) else base
    elif cls == 10:                                       # array
        nd = buf[q]
        q += 4 if ver < 3 else 1
        dims = struct.unpack_from("<%dI" % nd, buf, q)
        q += 4 * nd * (2 if ver < 3 else 1)               # v2 carries permutation indices too
        base, used = decode_datatype(buf, q)
        q += used
        dt = np.dtype((base, tuple(int(d) for d in dims)))
    elif cls == 9:
        raise HDF5FormatError("variable-length datatypes are not supported")
    else:
        raise HDF5FormatError("datatype class %d is not supported" % cls)
    return dt, q - p


def _float_props(size):
    if size == 8:
        return struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if size == 4:
        return struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    raise NotImplementedError("float%d" % (8 * size))


def encode_datatype(dt):
    """Datatype message of a (little-endian) numpy dtype, as libhdf5 1.8 encodes it."""
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise NotImplementedError("big-endian data is not written")
    size = dt.itemsize
    if dt.subdtype is not None:                           # array datatype, version 2
        base, shape = dt.subdtype
        nd = len(shape)
        return (struct.pack("<B3BI", 0x2A, 0, 0, 0, size) + struct.pack("<B3x", nd)
                + struct.pack("<%dI" % nd, *shape) + struct.pack("<%dI" % nd, *range(nd)) + encode_datatype(base))
    if dt.kind == "f":
        return struct.pack("<B3BI", 0x11, 0x20, 8 * size - 1, 0, size) + _float_props(size)
    if dt.kind in "iu":
        return struct.pack("<B3BI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, size) + struct.pack("<HH", 0, 8 * size)
    if dt.kind == "b":                                    # PyTables' bool: the 8-bit bit field H5T_STD_B8
        return struct.pack("<B3BI", 0x14, 0, 0, 0, 1) + struct.pack("<HH", 0, 8)
    if dt.kind == "S":
        return struct.pack("<B3BI", 0x13, 0, 0, 0, max(size, 1))
    if dt.kind == "c":
        half = np.dtype("f%d" % (size // 2))
        dt = np.dtype([("r", half), ("i", half)])
    if dt.kind == "V" and dt.names:
        ver = 2 if any(dt.fields[n][0].subdtype is not None for n in dt.names) else 1
        out = struct.pack("<B3BI", (ver << 4) | 6, len(dt.names) & 0xFF, len(dt.names) >> 8, 0, size)
        for name in dt.names:
            mdt, off = dt.fields[name][0], dt.fields[name][1]
            raw = name.encode("utf-8")
            out += raw + b"\0" * (((len(raw) + 8) // 8) * 8 - len(raw)) + struct.pack("<I", off)
            if ver == 1:
                out += b"\0" * 28
            out += encode_datatype(mdt)
        return out
    raise NotImplementedError("numpy dtype %r has no HDF5 encoding here" % (dt,))


# ------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------
class AttributeSet(dict):
    """node.attrs / node._v_attrs: a dict with PyTables' attribute-style access."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def _f_list(self, attrset="user"):
        return sorted(self)


class _Image(object):
    """The mapped file plus the superblock fields every structure needs."""

    def __init__(self, filename):
        self.filename = filename
        self.fh = open(filename, "rb")
        size = os.fstat(self.fh.fileno()).st_size
        if size < 96:
            self.fh.close()
            raise HDF5FormatError("%s: not an HDF5 file (too short)" % filename)
        self.mm = mmap.mmap(self.fh.fileno(), 0, access=mmap.ACCESS_READ)
        off = 0
        while self.mm[off:off + 8] != SIGNATURE:          # user block: 512, 1024, 2048, ...
            off = 512 if off == 0 else off * 2
            if off + 8 > size:
                self.close()
                raise HDF5FormatError("%s: no HDF5 signature" % filename)
        self.sb = off
        ver = self.mm[off + 8]
        if ver > 1:
            self.close()
            raise HDF5FormatError("%s: superblock version %d (HDF5 1.10 'latest' format) is not supported"
                                  % (filename, ver))
        if self.mm[off + 13] != 8 or self.mm[off + 14] != 8:
            self.close()
            raise HDF5FormatError("only 8-byte offsets and lengths are supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", self.mm, off + 16)
        p = off + 24 + (4 if ver == 1 else 0)             # v1: indexed-storage K + 2 reserved bytes
        self.chunk_k = struct.unpack_from("<H", self.mm, off + 24)[0] if ver == 1 else _CHUNK_K
        self.base, _free, self.eof, _drv = struct.unpack_from("<4Q", self.mm, p)
        # the base address counts from the start of the file; in a file with a user block it is
        # the superblock's own offset (spec II.A: "base address")
        self.root_entry = p + 32
        if self.base == 0 and off:
            self.base = off

    def close(self):
        try:
            self.mm.close()
        except (BufferError, ValueError):                 # arrays still view the mapping
            pass
        self.fh.close()

    def at(self, addr):
        if addr == UNDEF:
            raise HDF5FormatError("undefined address dereferenced")
        return self.base + addr

    # ---- object headers (spec IV.A.1.a) -----------------------------------------------------
    def messages(self, addr):
        """[(type, flags, absolute offset of data, size)] of the version-1 object header at addr."""
        mm, p = self.mm, self.at(addr)
        if mm[p:p + 4] == b"OHDR":
            raise HDF5FormatError("version-2 object headers (HDF5 1.10 'latest' format) are not supported")
        if mm[p] != 1:
            raise HDF5FormatError("object header version %d at %#x" % (mm[p], p))
        nmsg, _ref, hsize = struct.unpack_from("<HII", mm, p + 2)
        chunks = [(p + 16, hsize)]
        out = []
        while chunks and len(out) < nmsg:
            q, n = chunks.pop(0)
            end = q + n
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", mm, q)
                out.append((mtype, mflags, q + 8, msize))
                if mtype == 0x10:                         # continuation block
                    caddr, clen = struct.unpack_from("<QQ", mm, q + 8)
                    chunks.append((self.at(caddr), clen))
                q += 8 + msize
        return out

    # ---- old-style groups (spec III.A.1, III.B, III.D) --------------------------------------
    def heap_data(self, addr):
        p = self.at(addr)
        if self.mm[p:p + 4] != b"HEAP":
            raise HDF5FormatError("local heap signature missing at %#x" % p)
        size, _free, data = struct.unpack_from("<QQQ", self.mm, p + 8)
        d = self.at(data)
        return self.mm[d:d + size]

    def group_entries(self, btree, heap):
        """[(name, object header address, cache type)] of a symbol-table group, in B-tree order."""
        names = self.heap_data(heap)
        out = []

        def walk(addr):
            p = self.at(addr)
            if self.mm[p:p + 4] != b"TREE" or self.mm[p + 4] != 0:
                raise HDF5FormatError("group B-tree node expected at %#x" % p)
            level, nent = self.mm[p + 5], struct.unpack_from("<H", self.mm, p + 6)[0]
            for i in range(nent):
                child = struct.unpack_from("<Q", self.mm, p + 24 + 8 + 16 * i)[0]
                if level:
                    walk(child)
                    continue
                s = self.at(child)
                if self.mm[s:s + 4] != b"SNOD":
                    raise HDF5FormatError("symbol table node expected at %#x" % s)
                for j in range(struct.unpack_from("<H", self.mm, s + 6)[0]):
                    noff, ohdr, cache = struct.unpack_from("<QQI", self.mm, s + 8 + 40 * j)
                    out.append((_cstring(names, noff)[0].decode("utf-8", "replace"), ohdr, cache))
        if btree != UNDEF:
            walk(btree)
        return out

    def chunk_index(self, addr, rank):
        """[(offsets tuple, stored bytes, filter mask, address)] of a version-1 chunk B-tree."""
        out = []
        if addr == UNDEF:
            return out
        ksize = 8 + 8 * (rank + 1)

        def walk(a):
            p = self.at(a)
            if self.mm[p:p + 4] != b"TREE" or self.mm[p + 4] != 1:
                raise HDF5FormatError("chunk B-tree node expected at %#x" % p)
            level, nent = self.mm[p + 5], struct.unpack_from("<H", self.mm, p + 6)[0]
            for i in range(nent):
                q = p + 24 + i * (ksize + 8)
                nbytes, mask = struct.unpack_from("<II", self.mm, q)
                offs = struct.unpack_from("<%dQ" % (rank + 1), self.mm, q + 8)
                child = struct.unpack_from("<Q", self.mm, q + ksize)[0]
                if level:
                    walk(child)
                else:
                    out.append((tuple(int(o) for o in offs[:rank]), nbytes, mask, child))
        walk(addr)
        return out


def _decode_dataspace(buf, p):
    ver, rank, flags = buf[p], buf[p + 1], buf[p + 2]
    if ver == 1:
        q = p + 8
    elif ver == 2:
        if buf[p + 3] == 2:                               # null dataspace
            return None, None
        q = p + 4
    else:
        raise HDF5FormatError("dataspace message version %d" % ver)
    dims = struct.unpack_from("<%dQ" % rank, buf, q)
    maxdims = struct.unpack_from("<%dQ" % rank, buf, q + 8 * rank) if flags & 1 else dims
    return tuple(int(d) for d in dims), tuple(None if d == UNDEF else int(d) for d in maxdims)


def _decode_filters(buf, p):
    ver, n = buf[p], buf[p + 1]
    out = []
    q = p + (8 if ver == 1 else 2)
    for _ in range(n):
        fid = struct.unpack_from("<H", buf, q)[0]
        if ver == 1 or fid >= 256:
            namelen = struct.unpack_from("<H", buf, q + 2)[0]
            q += 4
        else:
            namelen = 0
            q += 2
        flags, ncd = struct.unpack_from("<HH", buf, q)
        q += 4 + (_pad8(namelen) if ver == 1 else namelen)
        cd = struct.unpack_from("<%dI" % ncd, buf, q)
        q += 4 * ncd + (4 if (ver == 1 and ncd % 2) else 0)
        out.append((fid, flags, cd))
    return out


def _decode_attribute(buf, p):
    ver = buf[p]
    nsize, tsize, ssize = struct.unpack_from("<HHH", buf, p + 2)
    if ver == 1:
        q, pad = p + 8, _pad8
    elif ver in (2, 3):
        if buf[p + 1] & 3:
            raise HDF5FormatError("shared datatype / dataspace in an attribute")
        q, pad = p + (8 if ver == 2 else 9), (lambda n: n)
    else:
        raise HDF5FormatError("attribute message version %d" % ver)
    name = bytes(buf[q:q + nsize]).split(b"\0")[0].decode("utf-8", "replace")
    q += pad(nsize)
    dt, _ = decode_datatype(buf, q)
    q += pad(tsize)
    dims, _ = _decode_dataspace(buf, q)
    q += pad(ssize)
    if dims is None:
        return name, None
    count = int(np.prod(dims)) if dims else 1
    val = np.frombuffer(bytes(buf[q:q + count * dt.itemsize]), dtype=dt, count=count).reshape(dims)
    if dt.kind == "S":
        val = np.char.decode(val, "utf-8", "replace") if dims else val.reshape(())[()].decode("utf-8", "replace")
    elif not dims:
        val = val.reshape(())[()]
    return name, val


class Node(object):
    _v_name = "/"

    @property
    def attrs(self):
        return self._v_attrs

    def _f_getattr(self, name):
        return self._v_attrs[name]


class Group(Node):
    """A group of an open file: children by attribute (``g.yresult``), ``in``, iteration."""

    def __init__(self, name, title=""):
        self.__dict__["_v_name"] = name
        self.__dict__["_v_children"] = {}
        self.__dict__["_v_attrs"] = AttributeSet()
        if title is not None:
            self._v_attrs["TITLE"] = title

    def __getattr__(self, name):
        try:
            return self.__dict__["_v_children"][name]
        except KeyError:
            raise NoSuchNodeError("group %r does not have a child named %r" % (self._v_name, name))

    def __setattr__(self, name, value):
        raise AttributeError("use File.create_* to add nodes")

    def __contains__(self, name):
        return name in self._v_children

    def __iter__(self):
        return iter([self._v_children[k] for k in sorted(self._v_children)])

    def __len__(self):
        return len(self._v_children)

    _v_title = property(lambda self: self._v_attrs.get("TITLE", ""))

    def _f_get_child(self, name):
        """The child called ``name`` -- also when the name is not usable as an attribute
        (``__dict__``, ``_v_attrs``, ...: PyTables' natural-naming caveat)."""
        try:
            return self._v_children[name]
        except KeyError:
            raise NoSuchNodeError("group %r does not have a child named %r" % (self._v_name, name))

    _f_getChild = _f_get_child

    def __repr__(self):
        return "<Group %s: %s>" % (self._v_name, ", ".join(sorted(self._v_children)))


def _unshuffle(raw, itemsize):
    n = len(raw) // itemsize
    if itemsize <= 1 or n == 0:
        return raw
    body = np.frombuffer(raw, dtype=np.uint8, count=n * itemsize).reshape(itemsize, n).T
    return np.ascontiguousarray(body).tobytes() + raw[n * itemsize:]


def _shuffle(raw, itemsize):
    n = len(raw) // itemsize
    if itemsize <= 1 or n == 0:
        return raw
    body = np.frombuffer(raw, dtype=np.uint8, count=n * itemsize).reshape(n, itemsize).T
    return np.ascontiguousarray(body).tobytes() + raw[n * itemsize:]


class Leaf(Node):
    """A dataset.  Reading: ``leaf[...]`` / ``leaf.read()`` return NumPy arrays; a contiguous
    dataset is served from a memory map (``leaf.memmap()``), a chunked one chunk by chunk.
    Writing (files opened "w" / "a"): ``append`` for extendable arrays and tables."""

    def __init__(self, name, dtype, shape, kind, title="", maxshape=None):
        self._v_name, self.dtype, self.kind = name, np.dtype(dtype), kind
        self._shape = tuple(int(s) for s in shape)
        self.maxshape = tuple(maxshape) if maxshape is not None else self._shape
        self._v_attrs = AttributeSet()
        if title is not None:
            self._v_attrs["TITLE"] = title
        self._blocks = []            # arrays appended in this session (extendable kinds), in order
        self._stored = None          # (image, layout tuple) of the part already in a file
        self.filters = Filters()
        self.chunkshape = None
        self._row = None

    # ---- shape --------------------------------------------------------------------------------
    @property
    def shape(self):
        if self.kind in ("EARRAY", "TABLE"):
            return (self._stored_rows() + sum(b.shape[0] for b in self._blocks),) + self._shape[1:]
        return self._shape

    def _stored_rows(self):
        return self._stored[1]["shape"][0] if self._stored is not None else 0

    nrows = property(lambda self: self.shape[0] if self.shape else 1)
    ndim = property(lambda self: len(self.shape))
    size = property(lambda self: int(np.prod(self.shape)))
    title = property(lambda self: self._v_attrs.get("TITLE", ""))
    colnames = property(lambda self: list(self.dtype.names or ()))
    atom = property(lambda self: Atom(self.dtype))

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return "<%s %s %s %s>" % (self.kind, self._v_name, self.shape, self.dtype)

    # ---- reading ------------------------------------------------------------------------------
    def memmap(self):
        """np.memmap of a contiguous (or the bytes of a compact) stored dataset, else None."""
        if self._stored is None or self._blocks or self._stored[1]["class"] == 2:
            return None
        return self.memmap_stored()

    def _read_stored(self, lo, hi):
        """Rows/region [lo, hi) (per-axis bounds) of the stored part as a fresh array."""
        img, lay = self._stored
        shape = lay["shape"]
        if lay["class"] != 2:
            mm = self.memmap_stored()
            return np.array(mm[tuple(slice(a, b) for a, b in zip(lo, hi))])
        if img.mm.closed:
            raise ClosedFileError("dataset %s: the file %s is closed" % (self._v_name, img.filename))
        out = np.zeros(tuple(b - a for a, b in zip(lo, hi)), dtype=self.dtype)
        cshape = lay["chunk"]
        csize = int(np.prod(cshape)) * self.dtype.itemsize
        if "chunks" not in lay:                           # walk the chunk B-tree once per dataset
            lay["chunks"] = lay["index"]()
        for offs, nbytes, mask, addr in lay["chunks"]:
            c_lo = [max(o, a) for o, a in zip(offs, lo)]
            c_hi = [min(o + c, b, s) for o, c, b, s in zip(offs, cshape, hi, shape)]
            if any(a >= b for a, b in zip(c_lo, c_hi)):
                continue
            p = img.at(addr)
            raw = bytes(img.mm[p:p + nbytes])
            for i in range(len(lay["filters"]) - 1, -1, -1):
                fid, _flags, cd = lay["filters"][i]
                if mask & (1 << i):
                    continue
                if fid == FILTER_DEFLATE:
                    raw = zlib.decompress(raw)
                elif fid == FILTER_SHUFFLE:
                    raw = _unshuffle(raw, cd[0] if cd else self.dtype.itemsize)
                elif fid == FILTER_FLETCHER32:
                    raw = raw[:-4]
                else:
                    raise HDF5FormatError("dataset %s is stored with the %s filter (id %d), which is not "
                                          "part of HDF5 itself; repack it with zlib (ptrepack / h5repack)"
                                          % (self._v_name, _FILTER_NAMES.get(fid, "unknown"), fid))
            if len(raw) < csize:
                raise HDF5FormatError("short chunk in dataset %s" % self._v_name)
            chunk = np.frombuffer(raw, dtype=self.dtype, count=csize // self.dtype.itemsize).reshape(cshape)
            src = tuple(slice(a - o, b - o) for a, b, o in zip(c_lo, c_hi, offs))
            dst = tuple(slice(a - l, b - l) for a, b, l in zip(c_lo, c_hi, lo))
            out[dst] = chunk[src]
        return out

    def memmap_stored(self):
        img, lay = self._stored
        if lay["class"] == 1:
            if lay["addr"] == UNDEF or int(np.prod(lay["shape"])) == 0:
                return np.zeros(lay["shape"], dtype=self.dtype)
            return np.memmap(img.filename, dtype=self.dtype, mode="r", offset=img.at(lay["addr"]),
                             shape=lay["shape"])
        return np.frombuffer(lay["data"], dtype=self.dtype).reshape(lay["shape"])

    def _read_rows(self, lo, hi):
        """Region with bounds lo/hi per axis, across the stored part and this session's blocks."""
        shape = self.shape
        if not shape:
            return self.read()
        parts = []
        r0 = 0
        if self._stored is not None:
            n = self._stored_rows()
            a, b = max(lo[0], 0), min(hi[0], n)
            if a < b:
                parts.append(self._read_stored((a,) + tuple(lo[1:]), (b,) + tuple(hi[1:])))
            r0 = n
        rest = tuple(slice(a, b) for a, b in zip(lo[1:], hi[1:]))
        for blk in self._blocks:
            a, b = max(lo[0] - r0, 0), min(hi[0] - r0, blk.shape[0])
            if a < b:
                parts.append(np.asarray(blk[(slice(a, b),) + rest]))
            r0 += blk.shape[0]
        if not parts:
            return np.zeros(tuple(max(0, b - a) for a, b in zip(lo, hi)), dtype=self.dtype)
        return parts[0] if len(parts) == 1 else np.concatenate(parts, axis=0)

    def read(self, start=None, stop=None, step=None):
        shape = self.shape
        if not shape:
            if self._blocks:
                return np.array(self._blocks[0])
            return np.array(self.memmap_stored()) if self._stored[1]["class"] != 2 else \
                self._read_stored((), ())
        if start is None and stop is None and step is None:
            return self._read_rows((0,) * len(shape), shape)
        return self[slice(start, stop, step)]

    def __getitem__(self, key):
        shape = self.shape
        if not isinstance(key, tuple):
            key = (key,)
        if not shape:
            return self.read()[key if key != (Ellipsis,) else ()]
        if any(k is Ellipsis for k in key):
            i = [k is Ellipsis for k in key].index(True)
            key = key[:i] + (slice(None),) * (len(shape) - len(key) + 1) + key[i + 1:]
        if (len(key) > len(shape)
                or not all(isinstance(k, (int, np.integer, slice)) for k in key)):
            if isinstance(key[0], str) and self.dtype.names:       # table column
                return self.read()[key[0]]
            return self.read()[key]                       # fancy indexing: through the full array
        key = key + (slice(None),) * (len(shape) - len(key))
        lo, hi, local = [], [], []
        for k, n in zip(key, shape):
            if isinstance(k, slice):
                a, b, s = k.indices(n)
                if s > 0:
                    b = max(a, b)
                    lo.append(a); hi.append(b); local.append(slice(None, None, s))
                else:                                     # negative step: read the span, flip locally
                    cnt = len(range(a, b, s))
                    first = a + (cnt - 1) * s if cnt else a
                    lo.append(first if cnt else 0); hi.append(a + 1 if cnt else 0)
                    local.append(slice(None, None, s))
            else:
                i = int(k) + (n if k < 0 else 0)
                if not 0 <= i < n:
                    raise IndexError("index %d is out of bounds for axis of size %d" % (int(k), n))
                lo.append(i); hi.append(i + 1); local.append(0)
        return self._read_rows(tuple(lo), tuple(hi))[tuple(local)]

    def __iter__(self):
        data = self.read()
        return iter(data)

    def iterrows(self):
        return iter(self)

    def col(self, name):
        return self.read()[name]

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype)

    # ---- writing ------------------------------------------------------------------------------
    def append(self, rows):
        if self.kind not in ("EARRAY", "TABLE"):
            raise TypeError("%s nodes cannot be extended" % self.kind)
        if self.kind == "TABLE":
            if not (isinstance(rows, np.ndarray) and rows.dtype == self.dtype):
                rows = np.array([tuple(r) for r in rows], dtype=self.dtype)
            rows = rows.reshape(-1)
        else:
            rows = np.asarray(rows)
            if rows.dtype != self.dtype:
                rows = rows.astype(self.dtype)
            if rows.shape[1:] != self._shape[1:]:
                raise ValueError("the appended rows have shape %r, the array's rows %r"
                                 % (rows.shape[1:], self._shape[1:]))
        if rows.shape[0]:
            self._blocks.append(rows)

    @property
    def row(self):
        if self.kind != "TABLE":
            raise TypeError("only tables have rows")
        if self._row is None:
            self._row = _Row(self)
        return self._row

    def flush(self):
        pass


class _Row(object):
    """table.row: fill the fields, then ``append()`` (tables.tableextension.Row)."""

    def __init__(self, table):
        self._table = table
        self._rec = np.zeros((), dtype=table.dtype)

    def __setitem__(self, key, value):
        if key not in self._table.dtype.names:
            raise KeyError("no such column: %s" % (key,))
        self._rec[key] = value

    def __getitem__(self, key):
        return self._rec[key]

    def append(self):
        self._table.append(self._rec.reshape(1).copy())
        self._rec = np.zeros((), dtype=self._table.dtype)


def _load_tree(img):
    """Build the node tree of an open image."""

    def load(addr, name, depth=0):
        if depth > 64:
            raise HDF5FormatError("group nesting too deep (cycle?)")
        msgs = img.messages(addr)
        kinds = dict((m[0], m) for m in msgs)
        attrs = AttributeSet()
        for mtype, mflags, p, n in msgs:
            if mtype == 0x0C and not (mflags & 2):
                try:
                    an, av = _decode_attribute(img.mm, p)
                    attrs[an] = av
                except HDF5FormatError:
                    pass                                  # e.g. a variable-length string: not needed
        if 0x11 in kinds:                                 # symbol-table message: a group
            btree, heap = struct.unpack_from("<QQ", img.mm, kinds[0x11][2])
            g = Group(name, None)
            g.__dict__["_v_attrs"] = attrs
            for cname, caddr, cache in img.group_entries(btree, heap):
                if cache == 2:                            # soft link
                    continue
                g._v_children[cname] = load(caddr, cname, depth + 1)
            return g
        if 0x01 in kinds and 0x03 in kinds and 0x08 in kinds:
            if kinds[0x03][1] & 2:
                raise HDF5FormatError("dataset %s uses a committed (shared) datatype" % name)
            dt, _ = decode_datatype(img.mm, kinds[0x03][2])
            dims, maxdims = _decode_dataspace(img.mm, kinds[0x01][2])
            if dims is None:
                dims, maxdims = (0,), (0,)
            lay = _decode_layout(img, kinds[0x08][2], len(dims), dt)
            lay["shape"] = dims
            lay["filters"] = _decode_filters(img.mm, kinds[0x0B][2]) if 0x0B in kinds else []
            cls = attrs.get("CLASS")
            kind = cls if cls in ("ARRAY", "EARRAY", "CARRAY", "TABLE") else \
                ("TABLE" if dt.names and len(dims) == 1 and lay["class"] == 2 else
                 "EARRAY" if (lay["class"] == 2 and maxdims and maxdims[0] is None) else "ARRAY")
            if kind == "CARRAY":
                kind = "ARRAY"
            leaf = Leaf(name, dt, dims, kind, None, maxdims)
            leaf._v_attrs = attrs
            leaf._stored = (img, lay)
            if lay["class"] == 2:
                leaf.chunkshape = lay["chunk"]
                deflate = [f for f in lay["filters"] if f[0] == FILTER_DEFLATE]
                if deflate:                               # a rewrite ("a" mode) keeps the compression
                    leaf.filters = Filters(deflate[0][2][0] if deflate[0][2] else 1, "zlib",
                                           any(f[0] == FILTER_SHUFFLE for f in lay["filters"]))
            return leaf
        if 0x02 in kinds or 0x06 in kinds:
            raise HDF5FormatError("object %s uses link messages (HDF5 1.10 'latest' format)" % name)
        raise HDF5FormatError("object %s is neither a group nor a dataset" % name)

    return load(struct.unpack_from("<Q", img.mm, img.root_entry + 8)[0], "/")


def _decode_layout(img, p, rank, dt):
    buf = img.mm
    ver = buf[p]
    if ver in (1, 2):
        nd, cls = buf[p + 1], buf[p + 2]
        q = p + 8
        addr = UNDEF
        if cls != 0:
            addr = struct.unpack_from("<Q", buf, q)[0]
            q += 8
        dims = struct.unpack_from("<%dI" % nd, buf, q)
        q += 4 * nd
        if cls == 0:
            size = struct.unpack_from("<I", buf, q)[0]
            return {"class": 0, "data": bytes(buf[q + 4:q + 4 + size])}
        if cls == 1:
            return {"class": 1, "addr": addr}
        chunk = tuple(int(d) for d in dims[:rank])
        return {"class": 2, "chunk": chunk, "index": (lambda: img.chunk_index(addr, rank))}
    if ver in (3, 4):
        cls = buf[p + 1]
        if cls == 0:
            size = struct.unpack_from("<H", buf, p + 2)[0]
            return {"class": 0, "data": bytes(buf[p + 4:p + 4 + size])}
        if cls == 1:
            return {"class": 1, "addr": struct.unpack_from("<Q", buf, p + 2)[0]}
        if cls == 2 and ver == 3:
            nd = buf[p + 2]
            addr = struct.unpack_from("<Q", buf, p + 3)[0]
            dims = struct.unpack_from("<%dI" % nd, buf, p + 11)
            chunk = tuple(int(d) for d in dims[:rank])
            return {"class": 2, "chunk": chunk, "index": (lambda: img.chunk_index(addr, rank))}
    raise HDF5FormatError("data layout message version %d class %d is not supported" % (ver, buf[p + 1]))


# ------------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------------
class _Out(object):
    """Sequential writer that keeps every structure 8-byte aligned."""

    def __init__(self, fh):
        self.fh, self.pos = fh, 0

    def align(self):
        pad = (-self.pos) & 7
        if pad:
            self.fh.write(b"\0" * pad)
            self.pos += pad
        return self.pos

    def write(self, data):
        """Write bytes / a contiguous array at the next aligned address; returns that address."""
        addr = self.align()
        if isinstance(data, (bytes, bytearray)):
            mv = data
        else:                                             # complex / compound arrays have no "B" cast
            mv = memoryview(np.ascontiguousarray(data).reshape(-1).view(np.uint8))
        self.fh.write(mv)
        self.pos += len(mv)
        return addr


def _message(mtype, body, flags=0):
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    if len(body) > 0xFFF8:
        raise HDF5FormatError("object header message of %d bytes" % len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dataspace_message(shape, maxshape=None):
    rank = len(shape)
    body = struct.pack("<BBB5x", 1, rank, 1 if maxshape is not None else 0) + struct.pack("<%dQ" % rank, *shape)
    if maxshape is not None:
        body += struct.pack("<%dQ" % rank, *[UNDEF if m is None else m for m in maxshape])
    return body


def _attribute_message(name, value):
    if isinstance(value, str):
        raw = value.encode("utf-8")
        dt, shape, data = struct.pack("<B3BI", 0x13, 0, 0, 0, len(raw) + 1), (), raw + b"\0"
    elif isinstance(value, bytes):
        dt, shape, data = struct.pack("<B3BI", 0x13, 0, 0, 0, len(value) + 1), (), value + b"\0"
    else:
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf-8")
        if arr.dtype.kind == "O":
            raise NotImplementedError("attribute %s: object arrays cannot be stored" % name)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        dt, shape, data = encode_datatype(arr.dtype), arr.shape, arr.tobytes()       # tobytes: C order
    nm = name.encode("utf-8") + b"\0"
    ds = _dataspace_message(shape)
    body = (struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + nm + b"\0" * (_pad8(len(nm)) - len(nm))
            + dt + b"\0" * (_pad8(len(dt)) - len(dt)) + ds + b"\0" * (_pad8(len(ds)) - len(ds)) + data)
    return _message(0x0C, body)


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


def plan_chunkshape(shape, itemsize, target=_CHUNK_TARGET_BYTES):
    """Chunk shape of an extendable array: whole rows of the first axis, about ``target`` bytes;
    a single row larger than 1 GiB is split along its last axis (a chunk's byte count is a
    32-bit field)."""
    rest = [int(s) for s in shape[1:]]
    row = int(np.prod(rest)) * itemsize if rest else itemsize
    if row <= _CHUNK_MAX_BYTES:
        return (int(max(1, min(target // max(row, 1), 1 << 20))),) + tuple(rest)
    parts = -(-row // _CHUNK_MAX_BYTES)
    rest[-1] = -(-rest[-1] // parts)
    if int(np.prod(rest)) * itemsize > _CHUNK_MAX_BYTES:
        raise NotImplementedError("rows of %d bytes cannot be chunked along the last axis alone" % row)
    return (1,) + tuple(rest)


def _write_chunk_btree(out, entries, rank, chunkshape, itemsize):
    """Version-1 B-tree (node type 1) over ``entries`` = [(offsets, nbytes, address)] in row-major
    order of the offsets; returns the root address (an empty leaf for no chunks)."""
    ksize = 8 + 8 * (rank + 1)
    node_size = 24 + (2 * _CHUNK_K + 1) * ksize + 2 * _CHUNK_K * 8

    def key(offs, nbytes):
        return struct.pack("<II", nbytes, 0) + struct.pack("<%dQ" % (rank + 1), *(tuple(offs) + (0,)))

    if entries:
        last = tuple(o + c for o, c in zip(entries[-1][0], chunkshape))
        end_key = struct.pack("<II", 0, 0) + struct.pack("<%dQ" % (rank + 1), *(last + (itemsize,)))
    else:
        end_key = b"\0" * ksize
    level = 0
    items = [(key(o, n), a) for o, n, a in entries]       # (first key of the subtree, address)
    while True:
        groups = [items[i:i + 2 * _CHUNK_K] for i in range(0, len(items), 2 * _CHUNK_K)] or [[]]
        base = out.align()
        addrs = [base + i * _pad8(node_size) for i in range(len(groups))]
        nxt = []
        for gi, grp in enumerate(groups):
            node = struct.pack("<4sBBHQQ", b"TREE", 1, level, len(grp),
                               addrs[gi - 1] if gi else UNDEF,
                               addrs[gi + 1] if gi + 1 < len(groups) else UNDEF)
            for k, a in grp:
                node += k + struct.pack("<Q", a)
            # the closing key: the first key of the right neighbour, or one chunk past the end
            node += groups[gi + 1][0][0] if gi + 1 < len(groups) else end_key
            node += b"\0" * (node_size - len(node))
            got = out.write(node)
            assert got == addrs[gi]
            nxt.append((grp[0][0] if grp else end_key, addrs[gi]))
        if len(groups) == 1:
            return addrs[0]
        items, level = nxt, level + 1


def _iter_row_groups(blocks, rows_per_chunk):
    """Consecutive groups of ``rows_per_chunk`` rows over a list of row blocks (the last group
    may be short); a group inside one block is a view."""
    pending, have = [], 0
    for blk in blocks:
        i, n = 0, blk.shape[0]
        while i < n:
            take = min(n - i, rows_per_chunk - have)
            pending.append(blk[i:i + take])
            have += take
            i += take
            if have == rows_per_chunk:
                yield pending[0] if len(pending) == 1 else np.concatenate(pending, axis=0)
                pending, have = [], 0
    if have:
        yield pending[0] if len(pending) == 1 else np.concatenate(pending, axis=0)


def _write_leaf(out, leaf):
    dt = leaf.dtype
    shape = leaf.shape
    pytables = {"ARRAY": ("ARRAY", "2.4"), "EARRAY": ("EARRAY", "1.1"), "TABLE": ("TABLE", "2.7")}[leaf.kind]
    attrs = AttributeSet(leaf._v_attrs)
    attrs.setdefault("TITLE", "")
    attrs["CLASS"], attrs["VERSION"] = pytables
    msgs = []
    if leaf.kind == "ARRAY":
        attrs.setdefault("FLAVOR", "numpy")
        data = np.ascontiguousarray(leaf.read(), dtype=dt)
        addr = out.write(data) if data.size else UNDEF
        msgs.append(_message(0x01, _dataspace_message(shape)))
        msgs.append(_message(0x03, encode_datatype(dt), 1))
        msgs.append(_message(0x05, struct.pack("<BBBBI", 2, 2, 2, 1, 0), 1))     # default fill value, late allocation
        msgs.append(_message(0x08, struct.pack("<BBQQ", 3, 1, addr, data.nbytes)))
    else:
        if leaf.kind == "EARRAY":
            attrs["EXTDIM"] = np.int32(0)
        else:
            attrs["NROWS"] = np.int64(shape[0])
            for i, name in enumerate(dt.names):
                attrs["FIELD_%d_NAME" % i] = name
                fdt = dt.fields[name][0]
                if fdt.subdtype is None:
                    attrs["FIELD_%d_FILL" % i] = np.zeros((), dtype=fdt)[()] if fdt.kind != "S" else b""
        chunk = tuple(leaf.chunkshape) if leaf.chunkshape else plan_chunkshape(shape, dt.itemsize)
        rank = len(shape)
        if len(chunk) != rank or any(c < 1 for c in chunk):
            raise ValueError("chunkshape %r does not fit an array of rank %d" % (chunk, rank))
        level = leaf.filters.complevel
        pipeline = []
        if level:
            if leaf.filters.shuffle:
                pipeline.append((FILTER_SHUFFLE, b"shuffle\0", (dt.itemsize,)))
            pipeline.append((FILTER_DEFLATE, b"deflate\0", (level,)))
        import itertools
        rows = _row_blocks(leaf, max(1, chunk[0]) * 64)
        entries = []
        r0 = 0
        grid_rest = [range(0, s, c) for s, c in zip(shape[1:], chunk[1:])]
        full_rest = all(c >= s for s, c in zip(shape[1:], chunk[1:]))
        for grp in _iter_row_groups(rows, chunk[0]):
            for rest in (itertools.product(*grid_rest) if grid_rest else [()]):
                piece = grp if full_rest and chunk[1:] == shape[1:] else \
                    grp[(slice(None),) + tuple(slice(o, o + c) for o, c in zip(rest, chunk[1:]))]
                if piece.shape != chunk:                  # edge chunks are stored at full size
                    full = np.zeros(chunk, dtype=dt)
                    full[tuple(slice(0, s) for s in piece.shape)] = piece
                    piece = full
                piece = np.ascontiguousarray(piece)
                if pipeline:
                    raw = piece.tobytes()
                    if leaf.filters.shuffle:
                        raw = _shuffle(raw, dt.itemsize)
                    raw = zlib.compress(raw, level)
                    addr, nbytes = out.write(raw), len(raw)
                else:
                    addr, nbytes = out.write(piece), piece.nbytes
                entries.append(((r0,) + tuple(rest), nbytes, addr))
            r0 += grp.shape[0]
        if r0 != shape[0]:
            raise HDF5FormatError("internal: wrote %d of %d rows of %s" % (r0, shape[0], leaf._v_name))
        btree = _write_chunk_btree(out, entries, rank, chunk, dt.itemsize)
        msgs.append(_message(0x01, _dataspace_message(shape, (None,) + tuple(shape[1:]))))
        msgs.append(_message(0x03, encode_datatype(dt), 1))
        msgs.append(_message(0x05, struct.pack("<BBBBI", 2, 3, 2, 1, 0), 1))     # default fill value, incremental allocation
        if pipeline:
            body = struct.pack("<BB6x", 1, len(pipeline))
            for fid, fname, cd in pipeline:
                body += struct.pack("<HHHH", fid, len(fname), 1, len(cd)) + fname
                body += struct.pack("<%dI" % len(cd), *cd) + (b"\0" * 4 if len(cd) % 2 else b"")
            msgs.append(_message(0x0B, body, 1))
        msgs.append(_message(0x08, struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", btree)
                             + struct.pack("<%dI" % (rank + 1), *(chunk + (dt.itemsize,)))))
    msgs.append(_message(0x12, struct.pack("<B3xI", 1, int(time.time()) & 0xFFFFFFFF)))
    for name in attrs:
        if attrs[name] is not None:
            msgs.append(_attribute_message(name, attrs[name]))
    return out.write(_object_header(msgs))


def _row_blocks(leaf, step):
    """Row blocks of an extendable leaf in order: the part stored in the file that is being
    replaced (``step`` rows at a time), then the arrays appended in this session."""
    if leaf._stored is not None:
        shape = leaf._stored[1]["shape"]
        for a in range(0, shape[0], step):
            yield leaf._read_stored((a,) + (0,) * (len(shape) - 1), (min(shape[0], a + step),) + tuple(shape[1:]))
    for blk in leaf._blocks:
        yield blk


def _write_group(out, group, is_root=False):
    """Children first, then local heap, symbol-table nodes, B-tree node and the object header.
    Returns (object header address, B-tree address, heap address)."""
    children = []
    for name in sorted(group._v_children, key=lambda s: s.encode("utf-8")):    # strcmp order
        node = group._v_children[name]
        if isinstance(node, Group):
            children.append((name,) + _write_group(out, node))
        else:
            children.append((name, _write_leaf(out, node), None, None))
    # local heap: offset 0 holds the empty string the first B-tree key points at
    heap = bytearray(8)
    name_off = []
    for name, _, _, _ in children:
        raw = name.encode("utf-8") + b"\0"
        name_off.append(len(heap))
        heap += raw + b"\0" * (_pad8(len(raw)) - len(raw))
    free_off = len(heap)
    seg = max(_pad8(free_off + 16), 88)                   # always leave one free block (>= 16 bytes)
    heap += struct.pack("<QQ", 1, seg - free_off) + b"\0" * (seg - free_off - 16)
    heap_addr = out.align()
    out.write(struct.pack("<4sB3xQQQ", b"HEAP", 0, seg, free_off, heap_addr + 32) + bytes(heap))
    # symbol-table nodes: between K and 2K entries each, filled evenly
    per = 2 * _GROUP_LEAF_K
    nnodes = max(1, -(-len(children) // per))
    if nnodes > 2 * _GROUP_INTERNAL_K:
        raise NotImplementedError("more than %d children in one group" % (per * 2 * _GROUP_INTERNAL_K))
    bounds = [(len(children) * i) // nnodes for i in range(nnodes + 1)]
    snods, keys = [], [0]
    for i in range(nnodes):
        part = list(range(bounds[i], bounds[i + 1]))
        body = struct.pack("<4sBxH", b"SNOD", 1, len(part))
        for j in part:
            name, ohdr, bt, hp = children[j]
            if bt is None:
                body += struct.pack("<QQII16x", name_off[j], ohdr, 0, 0)
            else:
                body += struct.pack("<QQIIQQ", name_off[j], ohdr, 1, 0, bt, hp)
        body += b"\0" * (8 + per * 40 - len(body))
        snods.append(out.write(body))
        keys.append(name_off[part[-1]] if part else 0)
    nent = nnodes if children else 0
    node = struct.pack("<4sBBHQQ", b"TREE", 0, 0, nent, UNDEF, UNDEF) + struct.pack("<Q", keys[0])
    for i in range(nent):
        node += struct.pack("<QQ", snods[i], keys[i + 1])
    node += b"\0" * (24 + (2 * _GROUP_INTERNAL_K + 1) * 8 + 2 * _GROUP_INTERNAL_K * 8 - len(node))
    btree_addr = out.write(node)
    attrs = AttributeSet(group._v_attrs)
    attrs.setdefault("TITLE", "")
    attrs["CLASS"], attrs["VERSION"] = "GROUP", "1.0"
    if is_root:
        attrs["PYTABLES_FORMAT_VERSION"] = "2.1"
    msgs = [_message(0x11, struct.pack("<QQ", btree_addr, heap_addr), 1)]
    for name in attrs:
        if attrs[name] is not None:
            msgs.append(_attribute_message(name, attrs[name]))
    return out.write(_object_header(msgs)), btree_addr, heap_addr


def _write_file(filename, root):
    with open(filename, "wb") as fh:
        out = _Out(fh)
        out.write(b"\0" * 96)                             # superblock, filled in at the end
        ohdr, btree, heap = _write_group(out, root, True)
        eof = out.align()
        sb = (SIGNATURE + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, _GROUP_LEAF_K, _GROUP_INTERNAL_K, 0)
              + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
              + struct.pack("<QQIIQQ", 0, ohdr, 1, 0, btree, heap))
        assert len(sb) == 96
        fh.seek(0)
        fh.write(sb)


# ------------------------------------------------------------------------------------------------
# File
# ------------------------------------------------------------------------------------------------
class File(object):
    """tables.File: ``mode`` "r" (read), "w" (new file, written at ``close``), "a" (the existing
    tree is loaded, nodes can be added / extended, and ``close`` writes a new file in its place)."""

    def __init__(self, filename, mode="r", title=""):
        if mode not in ("r", "w", "a", "r+"):
            raise ValueError("invalid mode %r" % (mode,))
        self.filename, self.mode = filename, ("a" if mode == "r+" else mode)
        self._img = None
        self.isopen = True
        if self.mode == "a" and not os.path.isfile(filename):
            self.mode = "w"
        if self.mode == "w":
            d = os.path.dirname(os.path.abspath(filename))
            if not os.path.isdir(d):
                raise IOError("Directory %s does not exist" % d)
            open(filename, "wb").close()                  # fail now, not at close, if it cannot be written
            self.root = Group("/", title)
        else:
            self._img = _Image(filename)
            try:
                self.root = _load_tree(self._img)
            except Exception:
                self._img.close()
                raise
        self._dirty = self.mode == "w"

    # ---- navigation ---------------------------------------------------------------------------
    def _where(self, where):
        self._check_open()
        if isinstance(where, Node):
            return where
        node = self.root
        for part in [p for p in str(where).split("/") if p]:
            if not isinstance(node, Group):
                raise NoSuchNodeError(where)
            node = node._f_get_child(part)
        return node

    def get_node(self, where, name=None):
        node = self._where(where)
        if name is None:
            return node
        if not isinstance(node, Group):
            raise NoSuchNodeError("%r is not a group" % (where,))
        return node._f_get_child(name)

    def list_nodes(self, where):
        return list(self._where(where))

    def walk_nodes(self, where="/"):
        stack = [self._where(where)]
        while stack:
            node = stack.pop(0)
            yield node
            if isinstance(node, Group):
                stack = list(node) + stack

    def __contains__(self, path):
        try:
            self._where(path)
            return True
        except NoSuchNodeError:
            return False

    def _check_open(self):
        if not self.isopen:
            raise ClosedFileError("the file %s is closed" % self.filename)

    # ---- creation -----------------------------------------------------------------------------
    def _add(self, where, name, node):
        self._check_open()
        if self.mode == "r":
            raise IOError("the file %s is opened read-only" % self.filename)
        parent = self._where(where)
        if not isinstance(parent, Group):
            raise TypeError("%r is not a group" % (where,))
        if not name or "/" in name:
            raise ValueError("invalid node name %r" % (name,))
        if name in parent._v_children:
            raise NodeError("group %s already has a child named %s" % (parent._v_name, name))
        parent._v_children[name] = node
        self._dirty = True
        return node

    def create_group(self, where, name, title="", filters=None):
        return self._add(where, name, Group(name, title))

    def create_array(self, where, name, obj=None, title="", atom=None, shape=None):
        if obj is None:
            obj = np.zeros(shape, dtype=atom.dtype)
        arr = np.asarray(obj)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf-8")
        if arr.dtype.kind == "O":
            raise TypeError("object arrays cannot be stored")
        leaf = Leaf(name, arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype,
                    arr.shape, "ARRAY", title)
        leaf._blocks = [arr]
        return self._add(where, name, leaf)

    def create_earray(self, where, name, atom=None, shape=None, title="", filters=None,
                      expectedrows=1000, chunkshape=None, obj=None):
        if atom is None:
            atom = Atom(np.asarray(obj).dtype)
        if shape is None:
            shape = (0,) + tuple(np.asarray(obj).shape[1:])
        shape = tuple(int(s) for s in shape)
        if not shape or shape[0] != 0:
            raise ValueError("the extendable dimension must be the first one and have length 0 "
                             "(the reference's files are laid out that way)")
        leaf = Leaf(name, atom.dtype, shape, "EARRAY", title, (None,) + shape[1:])
        leaf.filters = filters or Filters()
        leaf.chunkshape = tuple(chunkshape) if chunkshape else None
        if obj is not None:
            leaf.append(obj)
        return self._add(where, name, leaf)

    def create_table(self, where, name, description, title="", filters=None, expectedrows=10000,
                     chunkshape=None):
        if isinstance(description, dict):                 # PyTables orders a dict's columns by name
            description = np.dtype([(k, description[k].dtype if isinstance(description[k], Atom)
                                     else np.dtype(description[k])) for k in sorted(description)])
        dt = np.dtype(description)
        if not dt.names:
            raise TypeError("a table needs a compound description")
        leaf = Leaf(name, dt, (0,), "TABLE", title, (None,))
        leaf.filters = filters or Filters()
        leaf.chunkshape = tuple(chunkshape) if chunkshape else (max(1, min(1 << 14, (1 << 16) // dt.itemsize)),)
        return self._add(where, name, leaf)

    def remove_node(self, where, name=None, recursive=False):
        parent, node = (self._where(where), None)
        if name is None:
            raise ValueError("remove_node needs the parent and the child's name")
        if name not in parent._v_children:
            raise NoSuchNodeError(name)
        del parent._v_children[name]
        self._dirty = True

    def copy_node(self, node, newparent, newname=None, recursive=False):
        """Copy ``node`` (of this or of another open file) under ``newparent`` of this file.  A
        dataset stored in the other file is not read here: its bytes are fetched when this file
        is written, so the source must stay open until this file is closed."""
        name = newname or node._v_name
        if isinstance(node, Group):
            g = self.create_group(newparent, name, node._v_attrs.get("TITLE", ""))
            if recursive:
                for child in node:
                    self.copy_node(child, g, recursive=True)
            return g
        leaf = Leaf(name, node.dtype, node.shape if node.kind == "ARRAY" else (0,) + tuple(node.shape[1:]),
                    node.kind, node._v_attrs.get("TITLE", ""), node.maxshape)
        leaf.filters, leaf.chunkshape = node.filters, node.chunkshape
        leaf._stored, leaf._blocks = node._stored, list(node._blocks)
        return self._add(newparent, name, leaf)

    # PyTables 2.x spellings, which the reference uses
    copyNode = copy_node
    createGroup, createArray, createEArray, createTable = create_group, create_array, create_earray, create_table
    getNode, listNodes, walkNodes, removeNode = get_node, list_nodes, walk_nodes, remove_node

    # ---- life cycle ---------------------------------------------------------------------------
    def flush(self):
        self._check_open()

    def _modified(self):
        if self._dirty:
            return True
        return any(isinstance(n, Leaf) and n._stored is not None and n._blocks for n in self.walk_nodes())

    def close(self):
        if not self.isopen:
            return
        try:
            if self.mode != "r" and self._modified():
                tmp = self.filename + ".tmp%d" % os.getpid()
                try:
                    _write_file(tmp, self.root)
                except Exception:
                    if os.path.exists(tmp):
                        os.remove(tmp)
                    raise
                if self._img is not None:
                    self._img.close()
                    self._img = None
                os.replace(tmp, self.filename)
        finally:
            self.isopen = False
            if self._img is not None:
                self._img.close()
                self._img = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            if self.isopen and self.mode == "r":
                self.close()
        except Exception:
            pass


def open_file(filename, mode="r", title="", **kwargs):
    return File(filename, mode, title)


openFile = open_file


def is_hdf5_file(filename):
    """True when ``filename`` starts with the HDF5 signature (at 0 or after a user block)."""
    try:
        with open(filename, "rb") as fh:
            off = 0
            size = os.fstat(fh.fileno()).st_size
            while off + 8 <= size:
                fh.seek(off)
                if fh.read(8) == SIGNATURE:
                    return True
                off = 512 if off == 0 else off * 2
    except IOError:
        pass
    return False


isHDF5File = is_hdf5_file
