"""Minimal HDF5 reader (this image has no h5py): version-0 superblock, old-style groups (symbol-table B-trees + local heaps),
version-1 object headers, contiguous little-endian float / integer datasets.  Enough for the Keras weight file of the reference's
joint-limit classifier (assets/realistic_arm_limits_model.h5, loaded with keras.models.load_model in envs/env.py:39); used by
tools/compile_assets.py only."""
import struct
import numpy as np


class MiniH5:
    def __init__(self, path):
        self.b = open(path, 'rb').read()
        b = self.b
        assert b[:8] == b'\x89HDF\r\n\x1a\n'
        ver = b[8]
        assert ver == 0, 'superblock version %d' % ver
        self.so, self.sl = b[13], b[14]          # size of offsets / lengths
        assert self.so == 8 and self.sl == 8
        # superblock v0: 8 sig, 1 ver, 1 fs ver, 1 root ver, 1 res, 1 shm ver, 1 so, 1 sl, 1 res, 2 leaf k, 2 internal k, 4 flags, 4x8 addresses, then root symbol table entry
        p = 8 + 8 + 2 + 2 + 4 + 4 * 8
        self.root = self._ste(p)

    def _u(self, p, n):
        return int.from_bytes(self.b[p:p + n], 'little')

    def _ste(self, p):
        """symbol table entry: link name offset, object header address, cache type, reserved, scratch (btree, heap)"""
        name_off, ohdr, cache = self._u(p, 8), self._u(p + 8, 8), self._u(p + 16, 4)
        bt = hp = None
        if cache == 1:
            bt, hp = self._u(p + 24, 8), self._u(p + 32, 8)
        return dict(name_off=name_off, ohdr=ohdr, cache=cache, btree=bt, heap=hp)

    def _heap_data(self, addr):
        assert self.b[addr:addr + 4] == b'HEAP'
        return self._u(addr + 8 + 8 + 8, 8)       # data segment address

    def _name(self, heap_addr, off):
        d = self._heap_data(heap_addr) + off
        e = self.b.index(b'\0', d)
        return self.b[d:e].decode()

    def _btree_entries(self, addr, heap):
        assert self.b[addr:addr + 4] == b'TREE', self.b[addr:addr + 4]
        ntype, level, used = self.b[addr + 4], self.b[addr + 5], self._u(addr + 6, 2)
        assert ntype == 0
        p = addr + 8 + 16
        out = []
        # keys and children interleaved: key0 child0 key1 child1 ... keyN
        p += 8
        for i in range(used):
            child = self._u(p, 8); p += 8
            p += 8
            if level > 0:
                out += self._btree_entries(child, heap)
            else:
                assert self.b[child:child + 4] == b'SNOD'
                n = self._u(child + 6, 2)
                q = child + 8
                for j in range(n):
                    e = self._ste(q); q += 40
                    e['name'] = self._name(heap, e['name_off'])
                    out.append(e)
        return out

    def messages(self, ohdr):
        b = self.b
        assert b[ohdr] == 1, 'object header version %d' % b[ohdr]
        nmsg, size = self._u(ohdr + 2, 2), self._u(ohdr + 8, 4)
        blocks = [(ohdr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end and len(out) < nmsg:
                t, s, fl = self._u(p, 2), self._u(p + 2, 2), b[p + 4]
                body = p + 8
                if t == 0x10:                         # continuation
                    blocks.append((self._u(body, 8), self._u(body + 8, 8)))
                out.append((t, body, s))
                p = body + s
        return out

    def group(self, entry):
        """children of a group given its symbol table entry (or object header address)"""
        bt, hp = entry.get('btree'), entry.get('heap')
        if bt is None:
            for t, body, s in self.messages(entry['ohdr']):
                if t == 0x11:
                    bt, hp = self._u(body, 8), self._u(body + 8, 8)
        if bt is None:
            return {}
        return {e['name']: e for e in self._btree_entries(bt, hp)}

    def dataset(self, entry):
        shape = dtype = addr = size = None
        for t, body, s in self.messages(entry['ohdr']):
            b = self.b
            if t == 0x01:                             # dataspace v1
                ver, rank, flags = b[body], b[body + 1], b[body + 2]
                assert ver == 1
                shape = tuple(self._u(body + 8 + 8 * i, 8) for i in range(rank))
            elif t == 0x03:                           # datatype
                cls = b[body] & 0x0f
                sz = self._u(body + 4, 4)
                if cls == 1 and (b[body + 1] & 1) == 0:
                    dtype = {4: np.float32, 8: np.float64}[sz]
                elif cls == 0 and (b[body + 1] & 1) == 0:
                    dtype = {4: np.int32, 8: np.int64}[sz]
                else:
                    return None
            elif t == 0x08:                           # layout
                ver = b[body]
                if ver == 3:
                    assert b[body + 1] == 1, 'contiguous layout only (class %d)' % b[body + 1]
                    addr, size = self._u(body + 2, 8), self._u(body + 10, 8)
                else:
                    raise AssertionError('layout version %d' % ver)
        if dtype is None or addr is None:
            return None
        a = np.frombuffer(self.b, dtype=dtype, count=int(np.prod(shape)), offset=addr).reshape(shape).copy()
        return a

    def walk(self, entry=None, prefix=''):
        entry = entry or self.root
        out = {}
        for name, e in self.group(entry).items():
            path = prefix + '/' + name
            kids = None
            try:
                kids = self.group(e)
            except AssertionError:
                kids = None
            is_ds = any(t == 0x08 for t, _, _ in self.messages(e['ohdr']))
            if is_ds:
                out[path] = self.dataset(e)
            else:
                out.update(self.walk(e, path))
        return out

    def attr_strings(self, entry=None):
        """name -> raw bytes of every attribute of an object (version-1 attribute messages)"""
        entry = entry or self.root
        out = {}
        for t, body, s in self.messages(entry['ohdr']):
            if t != 0x0c:
                continue
            b = self.b
            ver = b[body]
            nsz, dtsz, dssz = self._u(body + 2, 2), self._u(body + 4, 2), self._u(body + 6, 2)
            pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
            p = body + 8
            name = b[p:p + nsz].split(b'\0')[0].decode(); p += pad(nsz)
            dt = p; p += pad(dtsz); p += pad(dssz)
            out[name] = (b[dt] & 0x0f, b[p:body + s])
        return out


if __name__ == '__main__':
    import sys
    h = MiniH5(sys.argv[1])
    for k, v in h.walk().items():
        print(k, None if v is None else (v.shape, v.dtype, float(np.abs(v).max())))
    for k, (cls, raw) in h.attr_strings().items():
        print('attr', k, cls, raw[:120])
