"""Boundary B2: the external-indexing wire protocol, played from a raw TCP socket exactly as
lantern_cli/tests/external_index_server_test.rs does (14-point lattice, M=12, efc=64, ef=32), against lb200_index_server."""
import os
import socket
import struct
import subprocess
import time

import numpy as np
import pytest

from test_golden import LATTICE, PQ_CODEBOOK, G, int_bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SERVER = os.path.join(ROOT, "lantern_b200", "lb200_index_server")
INIT, END, ERR = 0x13333337, 0x31333337, 0x37333337


@pytest.fixture(scope="module")
def server():
    port = 7998  # the port the reference's own test uses
    p = subprocess.Popen([SERVER, "--port", str(port), "--quiet"])
    for _ in range(100):
        try:
            socket.create_connection(("127.0.0.1", port), timeout=0.2).close()
            break
        except OSError:
            time.sleep(0.1)
    # the probe connection above is a failed session on the server side (it answers with an error frame): fine
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


def session(port, params, rows, codebook=None, row_bytes=None):
    s = socket.create_connection(("127.0.0.1", port), timeout=60)
    version, kind = struct.unpack("<II", recv_exact(s, 8))
    assert version == 1 and kind == 1  # external_index_server_test.rs:147-166
    s.sendall(struct.pack("<12I", INIT, *params))
    if codebook is not None:
        for row in codebook:
            s.sendall(np.asarray(row, np.float32).tobytes())
        s.sendall(struct.pack("<I", END))
    status = recv_exact(s, 1)
    if status != b"\x00":
        rest = status + recv_exact(s, 7)
        assert struct.unpack("<I", rest[:4])[0] == ERR
        n, = struct.unpack("<I", rest[4:])
        return None, recv_exact(s, n).decode()
    for label, vec in rows:
        s.sendall(struct.pack("<Q", label) + vec)
    s.sendall(struct.pack("<I", END))
    head = recv_exact(s, 8)
    if struct.unpack("<I", head[:4])[0] == ERR and len(rows) == 0:
        n, = struct.unpack("<I", head[4:])
        return None, recv_exact(s, n).decode()
    count, = struct.unpack("<Q", head)
    size, = struct.unpack("<Q", recv_exact(s, 8))
    data = recv_exact(s, size)
    s.close()
    return count, np.frombuffer(data, np.uint8)


def test_indexing_f32_lattice(server, eng, port):
    # params: pq, metric(l2sq=3), quantization(f32=1), dim, m, efc, ef, num_centroids, num_subvectors, capacity, element_bits
    rows = [(i, v.astype(np.float32).tobytes()) for i, v in enumerate(LATTICE)]
    count, data = session(server, (0, 3, 1, 3, 12, 64, 32, 0, 0, 14, 32), rows)
    assert count == 14  # what the reference test asserts (index.size() equality, :316)
    # the returned bytes are a usearch/lantern index file: the oracle loads it and answers like the reference's own graph
    p = port.PortIndex(3, "l2sq", "f32", M=12, efc=64, ef=32)
    p.reserve(14)
    p.load_buffer(data)
    assert p.size() == 14
    for i, v in enumerate(LATTICE):
        k, d, _ = p.search(v, 1)
        assert k[0] == i and d[0] == 0
    assert len(data) == len(G["lattice_f32_file"])  # same layout/size as the reference's file for these 14 vectors


def test_indexing_hamming_bits(server, port):
    bits = int_bits(LATTICE)  # 3 x int32 = 96 bits per row (external_index_server_test.rs:587-602)
    rows = [(i, bits[i].tobytes()) for i in range(14)]
    count, data = session(server, (0, 8, 5, 96, 12, 64, 32, 0, 0, 14, 1), rows)
    assert count == 14
    p = port.PortIndex(96, "hamming", "b1", M=12, efc=64, ef=32)
    p.reserve(14)
    p.load_buffer(data)
    k, d, _ = p.search(bits[5], 3)
    assert k[0] == 5 and d[0] == 0 and d[1] == 1


def test_indexing_pq(server):
    X = (LATTICE * 0.1).astype(np.float32)
    rows = [(i, v.tobytes()) for i, v in enumerate(X)]
    count, data = session(server, (1, 3, 1, 3, 12, 64, 32, 4, 3, 14, 32), rows, codebook=PQ_CODEBOOK)
    assert count == 14
    n, = struct.unpack_from("<Q", data.tobytes(), 80)
    assert n == 14


def test_capacity_growth_and_errors(server):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((300, 8)).astype(np.float32)
    rows = [(i + 1, v.tobytes()) for i, v in enumerate(X)]
    count, data = session(server, (0, 3, 1, 8, 8, 32, 16, 0, 0, 10, 32), rows)  # estimated capacity 10 < 300 rows
    assert count == 300
    s = socket.create_connection(("127.0.0.1", server), timeout=10)
    recv_exact(s, 8)
    s.sendall(struct.pack("<12I", 0xDEADBEEF, *([0] * 11)))  # not an init message
    head = recv_exact(s, 8)
    assert struct.unpack("<I", head[:4])[0] == ERR
    n, = struct.unpack("<I", head[4:])
    assert b"init" in recv_exact(s, n)
