"""BGZF (blocked gzip) writer/reader -- what `bgzip` produces and `samtools`/`tabix` expect.

The reference shells out to `bgzip -f` for its output VCF (phaser/phaser.py:1848-1853) and reads BAM
through samtools; here both sides are done in-process.  Format: SAM/BAM specification section 4.1 -- a
series of gzip members, each <= 64 KiB, carrying a 'BC' extra subfield with the block size, terminated
by an empty EOF block.
"""
import struct
import zlib

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
MAX_BLOCK = 0xff00


def compress_block(data: bytes, level=6) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = c.compress(data) + c.flush()
    bsize = len(body) + 25
    head = struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, bsize)
    return head + body + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def compress_all(data: bytes, level=6, threads=0):
    """All BGZF blocks of `data` (fixed MAX_BLOCK-byte cuts, what BGZFWriter emits), compressed on a thread pool
    (zlib releases the GIL).  Returns the list of compressed blocks in order; the EOF block is not included."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    mv = memoryview(data)
    cuts = [bytes(mv[o:o + MAX_BLOCK]) for o in range(0, len(mv), MAX_BLOCK)]
    n = threads or min(16, os.cpu_count() or 1)
    if n <= 1 or len(cuts) < 4:
        return [compress_block(c, level) for c in cuts]
    with ThreadPoolExecutor(max_workers=n) as ex:
        return list(ex.map(lambda c: compress_block(c, level), cuts))


class BGZFWriter:
    def __init__(self, path, level=6):
        self.f = open(path, "wb")
        self.buf = bytearray()
        self.level = level
        self.coffset = 0          # compressed bytes written so far = file offset of the block being filled
        self.block_sizes = []     # compressed size of every block emitted so far (the EOF block is not listed)

    def tell_virtual(self):
        """BGZF virtual file offset of the next byte written: block start << 16 | offset inside the block"""
        return (self.coffset << 16) | len(self.buf)

    def _emit(self, raw):
        blk = compress_block(raw, self.level)
        self.f.write(blk)
        self.coffset += len(blk)
        self.block_sizes.append(len(blk))

    def write(self, data):
        if isinstance(data, str):
            data = data.encode()
        self.buf += data
        while len(self.buf) >= MAX_BLOCK:
            self._emit(bytes(self.buf[:MAX_BLOCK]))
            del self.buf[:MAX_BLOCK]

    def close(self):
        if self.buf:
            self._emit(bytes(self.buf))
            self.buf = bytearray()
        self.f.write(EOF_BLOCK)
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_all(path) -> bytes:
    """Inflate a whole BGZF (or plain multi-member gzip) file."""
    out = []
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        d = zlib.decompressobj(31)
        out.append(d.decompress(data[pos:]))
        used = len(data) - pos - len(d.unused_data)
        if used <= 0:
            break
        pos += used
    return b"".join(out)


class VirtualReader:
    """Random access through BGZF virtual offsets (what an index points at)."""

    def __init__(self, path):
        with open(path, "rb") as f:
            self.data = f.read()
        self.block = -1; self.raw = b""; self.within = 0; self.next_block = 0

    def _load(self, coffset):
        if coffset >= len(self.data):
            self.block = coffset; self.raw = b""; self.next_block = coffset
            return
        bsize = struct.unpack_from("<H", self.data, coffset + 16)[0] + 1
        d = zlib.decompressobj(31)
        self.raw = d.decompress(self.data[coffset:coffset + bsize])
        self.block = coffset; self.next_block = coffset + bsize

    def seek(self, voffset):
        c, w = voffset >> 16, voffset & 0xffff
        if c != self.block:
            self._load(c)
        self.within = w

    def tell(self):
        if self.within >= len(self.raw) and self.raw:
            return self.next_block << 16
        return (self.block << 16) | self.within

    def readline(self):
        out = []
        while True:
            if self.within >= len(self.raw):
                if self.next_block >= len(self.data) or (self.block >= 0 and not self.raw):
                    break
                self._load(self.next_block); self.within = 0
                if not self.raw:
                    break
            i = self.raw.find(b"\n", self.within)
            if i < 0:
                out.append(self.raw[self.within:]); self.within = len(self.raw)
            else:
                out.append(self.raw[self.within:i + 1]); self.within = i + 1
                break
        return b"".join(out)
