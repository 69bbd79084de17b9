import os, socket, struct, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["BNET_NVL"] = "0"
from bagua_net_b200.utils.abi import NetPlugin
p = NetPlugin(8); p.init()
handle, lcomm = p.listen(0)
# the handle starts with a sockaddr: family(2) port(2, BE) addr(4)
fam, = struct.unpack("<H", handle[:2]); port, = struct.unpack(">H", handle[2:4]); ip = socket.inet_ntoa(handle[4:8])
print("listener", fam, ip, port)
junk = []
for payload in [b"", b"GET / HTTP/1.0\r\n\r\n", os.urandom(7), os.urandom(32), os.urandom(300)]:
    s = socket.create_connection((ip, port)); 
    if payload: s.sendall(payload)
    junk.append(s)
time.sleep(0.2)
assert p.accept(lcomm, poll=False) is None, "garbage produced a comm"
for s in junk[:2]: s.close()
# now a real connect must still get through
res = {}
def conn(): res["c"] = p.connect(handle)
t = threading.Thread(target=conn); t.start()
rc = p.accept(lcomm, timeout=10)
t.join()
print("accepted", bool(rc), p.transport_of(rc))
import numpy as np
a = np.arange(1000, dtype=np.uint8) ; b = np.zeros(1000, dtype=np.uint8)
mhs = p.reg_mr(res["c"], a.ctypes.data, 1000); mhr = p.reg_mr(rc, b.ctypes.data, 1000)
rr = p.irecv(rc, b.ctypes.data, 1000, mhr); sr = p.isend(res["c"], a.ctypes.data, 1000, mhs)
assert p.wait(sr) == 1000 and p.wait(rr) == 1000 and (a == b).all()
print("ok after garbage")
