"""Boundary B2 without a GPU: lb200_index_server's handling of the frames that never reach the device -- the hello, a wrong
first message, init fields that would size buffers from nonsense, an oversized codebook -- plus, on a box without a CUDA
device, the library's "no CPU fallback" error arriving as an ERR_MSG frame (server.rs:563-573)."""
import os
import socket
import struct
import subprocess
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SERVER = os.path.join(ROOT, "lantern_b200", "lb200_index_server")
INIT, END, ERR = 0x13333337, 0x31333337, 0x37333337


@pytest.fixture(scope="module")
def server():
    port = 7941
    p = subprocess.Popen([SERVER, "--port", str(port), "--quiet"])
    for _ in range(100):
        try:
            socket.create_connection(("127.0.0.1", port), timeout=0.2).close()
            break
        except OSError:
            time.sleep(0.1)
    yield port
    p.terminate()
    p.wait(timeout=10)


def recv_exact(s, n):
    out = b""
    while len(out) < n:
        chunk = s.recv(n - len(out))
        if not chunk:
            raise EOFError("closed after %d of %d bytes" % (len(out), n))
        out += chunk
    return out


def error_reply(port, payload):
    s = socket.create_connection(("127.0.0.1", port), timeout=30)
    assert struct.unpack("<II", recv_exact(s, 8)) == (1, 1)  # protocol version, server type (server.rs:182-183)
    s.sendall(payload)
    tag, n = struct.unpack("<II", recv_exact(s, 8))
    assert tag == ERR
    msg = recv_exact(s, n).decode()
    s.close()
    return msg


def init(*params):
    return struct.pack("<12I", INIT, *params)


def test_wrong_first_message(server):
    assert "init message" in error_reply(server, struct.pack("<12I", END, *([0] * 11)))


@pytest.mark.parametrize("params,needle", [
    ((0, 3, 9, 3, 12, 64, 32, 0, 0, 14, 32), "quantization"),
    ((0, 3, 1, 0, 12, 64, 32, 0, 0, 14, 32), "dimensions"),
    ((0, 3, 1, 1 << 30, 12, 64, 32, 0, 0, 14, 32), "dimensions"),
    ((0, 3, 1, 3, 12, 64, 32, 0, 0, 14, 16), "32-bit float rows"),
    ((1, 3, 1, 4, 12, 64, 32, 0, 2, 14, 32), "centroids"),
    ((1, 3, 1, 4, 12, 64, 32, 300, 2, 14, 32), "centroids"),
])
def test_init_fields_are_checked_before_anything_is_sized_from_them(server, params, needle):
    assert needle in error_reply(server, init(*params))


def test_codebook_longer_than_announced(server):
    rows = np.arange(12, dtype=np.float32).reshape(3, 4)  # 3 rows for num_centroids = 2
    msg = error_reply(server, init(1, 3, 1, 4, 12, 64, 32, 2, 2, 14, 32) + rows.tobytes() + struct.pack("<I", END))
    assert "more rows than num_centroids" in msg


def test_without_a_device_the_error_travels_as_a_frame(server):
    from lantern_b200 import api
    if api.device_count() > 0:
        pytest.skip("a GPU is present: the GPU suite plays the full conversations")
    assert "CUDA device unavailable" in error_reply(server, init(0, 3, 1, 3, 12, 64, 32, 0, 0, 14, 32))


def test_server_survives_all_of_the_above(server):
    s = socket.create_connection(("127.0.0.1", server), timeout=5)
    assert struct.unpack("<II", recv_exact(s, 8)) == (1, 1)
    s.close()
