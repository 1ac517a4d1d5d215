"""The scan-side batching daemon: many single-query clients -> GPU batches, with scan.c's streaming pattern."""
import os
import socket
import struct
import subprocess
import threading
import time

import numpy as np
import pytest

from util import structured

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DAEMON = os.path.join(ROOT, "lantern_b200", "lb200_search_daemon")
MAGIC_Q, MAGIC_S = 0x3151424C, 0x5351424C


def recv_exact(s, n):
    out = b""
    while len(out) < n:
        c = s.recv(n - len(out))
        if not c:
            raise EOFError
        out += c
    return out


def ask(s, vec, k, ef=0, cont=False):
    body = b"" if cont else vec.tobytes()
    s.sendall(struct.pack("<5I", MAGIC_Q, k, ef, 1 if cont else 0, len(body)) + body)
    status, found = struct.unpack("<II", recv_exact(s, 8))
    if status:
        raise RuntimeError(recv_exact(s, found).decode())
    keys = np.frombuffer(recv_exact(s, 8 * found), np.uint64)
    dists = np.frombuffer(recv_exact(s, 4 * found), np.float32)
    return keys, dists


@pytest.mark.parametrize("devices", [None, "group"])
def test_many_backends_are_batched(eng, tmp_path, devices):
    """devices = "group": the daemon serves the index from a row-sharded multi-GPU group (--devices; two ranks, on two GPUs when
    the box has them, else both on GPU 0): same answers as the direct single-GPU batch call."""
    n, d = 20000, 64
    X = structured(n, d, seed=5)
    g = eng.Index(d, "l2sq", "f32", M=16, efc=64, ef=48)
    g.reserve(n)
    g.add_batch(np.arange(1, n + 1, dtype=np.uint64), X)
    g.build()
    path = str(tmp_path / "idx.usearch")
    g.save(path)
    sock = str(tmp_path / "lb200.sock")
    p = subprocess.Popen([DAEMON, "--index", path, "--dim", str(d), "--m", "16", "--ef-construction", "64", "--ef", "48",
                          "--socket", sock, "--window-us", "2000"] +
                         (["--devices", "0,1" if eng.device_count() > 1 else "0,0", "--max-batch", "256"] if devices else []),
                         stderr=subprocess.DEVNULL)
    try:
        for _ in range(600):
            if os.path.exists(sock):
                break
            time.sleep(0.05)
        nclients, per = 64, 8
        Q = structured(nclients * per, d, seed=6)
        want_k, want_d, _ = g.search_batch(Q, 10)
        out = [None] * nclients

        def client(ci):
            s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            s.connect(sock)
            res = [ask(s, Q[ci * per + j], 10) for j in range(per)]
            # scan.c streaming on the last query: next 20 rows, no repeats
            k2, d2 = ask(s, None, 20, cont=True)
            out[ci] = (res, k2, d2)
            s.close()

        th = [threading.Thread(target=client, args=(i,)) for i in range(nclients)]
        [t.start() for t in th]
        [t.join() for t in th]
        for ci in range(nclients):
            res, k2, d2 = out[ci]
            for j in range(per):
                assert np.array_equal(res[j][0], want_k[ci * per + j])  # same answers as a direct batch call
            assert len(k2) == 20 and not (set(k2) & set(res[-1][0]))
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.connect(sock)
        s.sendall(struct.pack("<I", MAGIC_S))
        nreq, nbatch, largest = struct.unpack("<3Q", recv_exact(s, 24))
        assert nreq == nclients * (per + 1)
        assert largest > 1 and nbatch < nreq  # concurrent backends really were served together
        with pytest.raises(RuntimeError, match="wrong size"):
            ask(s, np.zeros(3, np.float32), 5)
    finally:
        p.terminate()
        p.wait(timeout=10)
