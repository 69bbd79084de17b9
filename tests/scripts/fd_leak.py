"""Open and close many connections inside one process; the number of open file descriptors, threads and /dev/shm
files must not grow (a plugin lives as long as the training job and NCCL re-creates communicators)."""
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

from bagua_net_b200.utils.abi import NetPlugin  # noqa: E402


def nfds():
    return len(os.listdir("/proc/self/fd"))


def nthreads():
    return len(os.listdir("/proc/self/task"))


def main():
    p = NetPlugin(8)
    p.init()
    a = np.arange(70000, dtype=np.uint8)
    b = np.zeros(70000, dtype=np.uint8)

    def cycle():
        handle, lcomm = p.listen(0)
        res = {}
        t = threading.Thread(target=lambda: res.setdefault("c", p.connect(handle)))
        t.start()
        rc = p.accept(lcomm, timeout=20)
        t.join()
        sc = res["c"]
        mhs, mhr = p.reg_mr(sc, a.ctypes.data, a.size), p.reg_mr(rc, b.ctypes.data, b.size)
        rr, sr = p.irecv(rc, b.ctypes.data, b.size, mhr), p.isend(sc, a.ctypes.data, a.size, mhs)
        assert p.wait(sr) == a.size and p.wait(rr) == a.size
        p.dereg_mr(sc, mhs)
        p.dereg_mr(rc, mhr)
        p.close_send(sc)
        p.close_recv(rc)
        p.close_listen(lcomm)
        return p

    for _ in range(10):       # warm-up: pools, loops and caches reach their steady size
        cycle()
    f0, t0, s0 = nfds(), nthreads(), len(os.listdir("/dev/shm"))
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 100):
        cycle()
    f1, t1, s1 = nfds(), nthreads(), len(os.listdir("/dev/shm"))
    print(f"fds {f0} -> {f1}, threads {t0} -> {t1}, shm files {s0} -> {s1}")
    assert f1 <= f0 + 2 and t1 <= t0 + 2 and s1 <= s0, "leak"
    print("no leak")


if __name__ == "__main__":
    main()
