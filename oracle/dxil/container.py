"""ORACLE TOOLING (test infrastructure, not product code): disassemble the DXIL of a compiled reference shader.

The reference ships its shaders compiled (`/root/reference/Bin/*.cso`: DXBC containers whose DXIL chunk is LLVM 3.7 bitcode).
Three of them (CSSHCubeMap / CSSHSum / CSSHNormalize) have no HLSL source in the tree at all. llvmlite (LLVM's own bitcode
reader, in this image) parses that bitcode once one incompatibility is removed: the DXIL data layout string says `i8:32`, which
current LLVM rejects ("i8 must be 8-bit aligned"). `patch_layout` rewrites those characters in place inside the bitstream
(same length, same VBR6 chunk count, located with the minimal bitstream walker in bitstream.py); nothing else is touched.

    python -m oracle.dxil.container /root/reference/Bin/CSVolumeCull.cso > /tmp/CSVolumeCull.ll

The disassembly is never committed: `make_golden.py` interprets it here (interp.py) and commits only input / output vectors."""
import struct
import sys

from .bitstream import walk


def extract(path):
    """The LLVM bitcode of the DXIL chunk of a DXBC container."""
    data = open(path, "rb").read()
    if data[:4] != b"DXBC":
        raise ValueError(f"{path}: not a DXBC container")
    n = struct.unpack_from("<I", data, 28)[0]
    for o in struct.unpack_from("<%dI" % n, data, 32):
        if data[o:o + 4] == b"DXIL":
            body = data[o + 8:o + 8 + struct.unpack_from("<I", data, o + 4)[0]]
            magic, _ver, bcoff, bcsize = struct.unpack_from("<4sIII", body, 8)
            if magic != b"DXIL":
                raise ValueError("bad DXIL program header")
            return bytearray(body[8 + bcoff:8 + bcoff + bcsize])
    raise ValueError(f"{path}: no DXIL chunk")


def _write_bits(buf, pos, n, v):
    for i in range(n):
        byte, off = (pos + i) >> 3, (pos + i) & 7
        buf[byte] = (buf[byte] & ~(1 << off)) | (((v >> i) & 1) << off)


def patch_layout(bc):
    for path, code, ops, poss, how in walk(bytes(bc)):
        if path == (8,) and code == 3:                       # MODULE_BLOCK / MODULE_CODE_DATALAYOUT
            if how != "unabbrev":
                raise ValueError("abbreviated data layout record")
            old = "".join(map(chr, ops))
            new = old.replace("i8:32", "i8:08").replace("i16:32", "i16:16")
            for a, b, p in zip(old, new, poss):
                if a != b:                                    # both characters are >= 32: two VBR6 chunks each
                    _write_bits(bc, p, 6, (ord(b) & 31) | 32)
                    _write_bits(bc, p + 6, 6, ord(b) >> 5)
            return new
    raise ValueError("no data layout record")


def disassemble(path):
    import llvmlite.binding as llvm
    bc = extract(path)
    patch_layout(bc)
    return str(llvm.parse_bitcode(bytes(bc)))


if __name__ == "__main__":
    sys.stdout.write(disassemble(sys.argv[1]))
