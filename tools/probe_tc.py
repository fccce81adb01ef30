import subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, "/root/repo")
from uno_b200 import integral_operators as ops
B, Ci, Co, idim, odim, modes = eval(sys.argv[1])
torch.manual_seed(0)
m = ops.SpectralConv2d_Uno(Ci, Co, *odim, *modes).cuda()
x = torch.randn(B, Ci, *idim, device="cuda", requires_grad=True)
y = m(x); torch.cuda.synchronize(); print("fwd ok", end=" ")
y.backward(torch.randn_like(y)); torch.cuda.synchronize(); print("bwd ok")
'''
shapes = ["(2, 3, 4, (40, 64), (30, 240), (6, 18))","(1, 2, 3, (64, 64), (17, 481), (5, 18))","(2, 2, 2, (32, 32), (50, 120), (8, 8))","(1, 3, 2, (32, 32), (33, 31), (4, 5))","(2, 2, 3, (32, 48), (9, 16), (3, 7))","(1, 2, 2, (64, 64), (300, 64), (20, 22))","(1, 1, 2, (96, 96), (130, 446), (9, 32))"]
for s in shapes:
    try:
        r = subprocess.run([sys.executable, "-c", code, s], capture_output=True, text=True, timeout=40)
        print(s, "->", r.stdout.strip(), r.stderr.strip()[-200:])
    except subprocess.TimeoutExpired as e:
        print(s, "-> TIMEOUT", (e.stdout or b"")[-100:])
