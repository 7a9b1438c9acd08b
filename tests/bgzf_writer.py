"""Test-only BGZF writer (what the reference's utils/make_bgzf.py produces; BGZF output is out of
the product's scope, the product only READS it).  Blocks of <= 65280 input bytes, raw deflate."""
import struct
import zlib

import numpy as np

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def write_bgzf(path, payload, block_size=65280, level=6):
    if not 1 <= int(block_size) <= 65280:
        raise ValueError("block_size must be in [1, 65280]")
    data = payload.tobytes() if isinstance(payload, np.ndarray) else bytes(payload)
    out = []
    for a in range(0, len(data), block_size):
        chunk = data[a:a + block_size]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        total = len(comp) + 26
        assert total <= 65536
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", total - 1) + comp
                   + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    out.append(EOF_BLOCK)
    with open(path, "wb") as fp:
        fp.write(b"".join(out))
