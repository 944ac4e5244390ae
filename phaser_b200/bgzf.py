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


class BGZFWriter:
    def __init__(self, path, level=6):
        self.f = open(path, "wb")
        self.buf = bytearray()
        self.level = level

    def write(self, data):
        if isinstance(data, str):
            data = data.encode()
        self.buf += data
        while len(self.buf) >= MAX_BLOCK:
            self.f.write(compress_block(bytes(self.buf[:MAX_BLOCK]), self.level))
            del self.buf[:MAX_BLOCK]

    def close(self):
        if self.buf:
            self.f.write(compress_block(bytes(self.buf), self.level))
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
