"""Restates the reference's own exact (integer) unit tests against the oracle primitives:
   - src/lib/bs_encode/encode_ut.cpp:138-176  (bisection call counts / final bit counts)
   - src/lib/bitstream/bitstream_ut.cpp:24-134 (MSB-first writer, sign handling)
so the restated primitives are pinned by every golden value the reference's tests hold for them."""
import ctypes

import numpy as np

import atde_testlib as tl


class OBits(ctypes.Structure):
    _fields_ = [("buf", ctypes.c_uint8 * 4096), ("size", ctypes.c_int), ("bits_used", ctypes.c_int)]


def read_bits(buf, pos, n):
    v = 0
    for i in range(n):
        byte = buf[(pos + i) // 8]
        v = (v << 1) | ((byte >> (7 - (pos + i) % 8)) & 1)
    return v


def test_bisection_call_counts():
    lib = tl.port_lib()
    calls, bits = ctypes.c_int(), ctypes.c_long()
    lib.obisect_selftest(1, ctypes.byref(calls), ctypes.byref(bits))
    assert (calls.value, bits.value) == (8, 1000)          # BsEncode.SimpleAlloc
    lib.obisect_selftest(2, ctypes.byref(calls), ctypes.byref(bits))
    assert (calls.value, bits.value) == (11, 993)          # BsEncode.NotExactAlloc


def test_bitstream_default_and_simple_write():
    lib = tl.port_lib()
    b = OBits()
    lib.obits_init(ctypes.byref(b))
    assert b.size == 0 and b.bits_used == 0                # TBitStream.DefaultConstructor
    lib.obits_write(ctypes.byref(b), 5, 3)                 # TBitStream.SimpleWriteRead
    assert b.bits_used == 3 and b.size == 1
    assert read_bits(b.buf, 0, 3) == 5


def test_bitstream_overlap_round_trips():
    lib = tl.port_lib()
    lib.omake_sign.restype = ctypes.c_int
    b = OBits()
    lib.obits_init(ctypes.byref(b))
    seq = [(101, 22), (212, 22), (323, 22), (0x1FFFFF, 21), (1, 1), (0, 7), (0x7F, 7), (3, 2)]
    for v, n in seq:
        lib.obits_write(ctypes.byref(b), v, n)
    pos = 0
    for v, n in seq:
        assert read_bits(b.buf, pos, n) == v
        pos += n
    assert b.bits_used == pos and b.size * 8 >= pos


def test_bitstream_sign():
    lib = tl.port_lib()
    lib.omake_sign.restype = ctypes.c_int
    b = OBits()
    lib.obits_init(ctypes.byref(b))
    for v, n in [(-2, 3), (-1, 3), (3, 3), (-4, 3), (-100, 8)]:
        lib.obits_write(ctypes.byref(b), ctypes.c_uint32(v & 0xffffffff), n)
    pos = 0
    for v, n in [(-2, 3), (-1, 3), (3, 3), (-4, 3), (-100, 8)]:
        raw = read_bits(b.buf, pos, n)
        assert lib.omake_sign(raw, n) == v                 # MakeSign, bitstream.h:27-31
        pos += n


def test_first_set_bit_and_relation_idx_tables():
    """GetFirstSetBit (src/util_ut.cpp:32-42): index of the highest set bit."""
    for i in range(1, 32):
        x = 1 << i
        assert x.bit_length() - 1 == i
