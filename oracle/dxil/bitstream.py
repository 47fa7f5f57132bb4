"""Minimal LLVM bitstream walker: yields (block path, code, ops, bit position of every op) for each record."""
class Reader:
    def __init__(self, data):
        self.d = data; self.pos = 0
    def read(self, n):
        v = 0; got = 0
        while got < n:
            byte = self.d[self.pos >> 3]; off = self.pos & 7
            take = min(8 - off, n - got)
            v |= ((byte >> off) & ((1 << take) - 1)) << got
            got += take; self.pos += take
        return v
    def vbr(self, n):
        v = 0; shift = 0
        while True:
            c = self.read(n)
            v |= (c & ((1 << (n - 1)) - 1)) << shift
            shift += n - 1
            if not (c >> (n - 1)): return v
    def align32(self):
        self.pos = (self.pos + 31) & ~31

CHAR6 = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789._"

def walk(data):
    r = Reader(data)
    assert r.read(32) == 0xdec04342
    blockinfo = {}
    out = []
    def read_abbrev_def():
        n = r.vbr(5); ops = []
        i = 0
        while i < n:
            if r.read(1): ops.append(("lit", r.vbr(8)))
            else:
                e = r.read(3)
                if e == 1: ops.append(("fixed", r.vbr(5)))
                elif e == 2: ops.append(("vbr", r.vbr(5)))
                elif e == 3: ops.append(("array",))
                elif e == 4: ops.append(("char6",))
                elif e == 5: ops.append(("blob",))
                else: raise ValueError(e)
            i += 1
        return ops
    def scalar(op):
        p = r.pos
        if op[0] == "lit": return op[1], p
        if op[0] == "fixed": return (r.read(op[1]) if op[1] else 0), p
        if op[0] == "vbr": return (r.vbr(op[1]) if op[1] else 0), p
        if op[0] == "char6": return ord(CHAR6[r.read(6)]), p
        raise ValueError(op)
    def block(path, bid, width):
        abbrevs = list(blockinfo.get(bid, []))
        cur_bid = None
        while True:
            aid = r.read(width)
            if aid == 0:
                r.align32(); return
            if aid == 1:
                nb = r.vbr(8); nw = r.vbr(4); r.align32(); r.read(32)
                block(path + [nb], nb, nw); continue
            if aid == 2:
                a = read_abbrev_def()
                if bid == 0: blockinfo.setdefault(cur_bid, []).append(a)
                else: abbrevs.append(a)
                continue
            if aid == 3:
                code = r.vbr(6); n = r.vbr(6); ops = []; poss = []
                for _ in range(n):
                    poss.append(r.pos); ops.append(r.vbr(6))
                if bid == 0 and code == 1: cur_bid = ops[0]
                out.append((tuple(path), code, ops, poss, "unabbrev")); continue
            a = abbrevs[aid - 4]
            vals = []; poss = []
            i = 0
            while i < len(a):
                op = a[i]
                if op[0] == "array":
                    n = r.vbr(6); el = a[i + 1]
                    for _ in range(n):
                        v, p = scalar(el); vals.append(v); poss.append(p)
                    i += 2; continue
                if op[0] == "blob":
                    n = r.vbr(6); r.align32(); p = r.pos
                    vals.append(bytes(r.d[(p >> 3):(p >> 3) + n])); poss.append(p); r.pos += n * 8; r.align32()
                    i += 1; continue
                v, p = scalar(op); vals.append(v); poss.append(p); i += 1
            out.append((tuple(path), vals[0], vals[1:], poss[1:], a))
    # top level: abbrev width 2
    while r.pos + 32 <= len(data) * 8:
        aid = r.read(2)
        if aid != 1: break
        nb = r.vbr(8); nw = r.vbr(4); r.align32(); r.read(32)
        block([nb], nb, nw)
    return out
