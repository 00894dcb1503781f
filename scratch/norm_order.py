"""Which summation order does torch's CUDA vector norm over a size-3 last dim use?  (bit-exact in-kernel view dir)"""
import torch
torch.manual_seed(0)
N = 3000000
xyz = ((torch.rand(N, 3, device="cuda") * 2 - 1) * 4.0)
cam = torch.tensor([5.1234, -1.777, 3.3339], device="cuda")
v = xyz - cam[None]
n_t = v.norm(dim=-1, keepdim=True)
x, y, z = v[:, 0], v[:, 1], v[:, 2]
sq = lambda a: a * a
d = lambda t: t.double()
def fma(a, b, c): return (d(a) * d(b) + d(c)).float()
cands = {
    "(x2+y2)+z2": (sq(x) + sq(y)) + sq(z),
    "(x2+z2)+y2": (sq(x) + sq(z)) + sq(y),
    "x2+(y2+z2)": sq(x) + (sq(y) + sq(z)),
    "fma(z,z,fma(y,y,x2))": fma(z, z, fma(y, y, sq(x))),
    "fma(y,y,fma(z,z,x2))": fma(y, y, fma(z, z, sq(x))),
    "fma(x,x,fma(y,y,z2))": fma(x, x, fma(y, y, sq(z))),
    "fma(z,z,x2)+y2": fma(z, z, sq(x)) + sq(y),
    "fma(y,y,x2)+z2": fma(y, y, sq(x)) + sq(z),
    "exact(double)": (d(x) ** 2 + d(y) ** 2 + d(z) ** 2).float(),
}
for k, s in cands.items():
    n = torch.sqrt(s)
    print("%-24s norm bit-equal %.6f   query bit-equal %.6f" % (k, (n == n_t[:, 0]).float().mean().item(),
          ((v / n[:, None]) == (v / n_t)).all(dim=1).float().mean().item()))
# rows by alignment class
s = (sq(x) + sq(y)) + sq(z)
eq = torch.sqrt(s) == n_t[:, 0]
for r in range(4):
    print("row %% 4 == %d: (x2+y2)+z2 matches %.6f" % (r, eq[r::4].float().mean().item()))
n2 = torch.linalg.vector_norm(v, dim=-1, keepdim=True)
print("vector_norm == norm:", torch.equal(n2, n_t))
