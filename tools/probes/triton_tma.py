import torch, triton, triton.language as tl
from triton.tools.tensor_descriptor import TensorDescriptor

@triton.jit
def k(desc, out_ptr, X0, Y0, BH: tl.constexpr, BW: tl.constexpr):
    t = desc.load([Y0, X0])
    r = tl.arange(0, BH)[:, None]; c = tl.arange(0, BW)[None, :]
    tl.store(out_ptr + r * BW + c, t)

x = torch.arange(320 * 640, device="cuda", dtype=torch.float32).reshape(320, 640)
d = TensorDescriptor.from_tensor(x, [16, 32])
out = torch.zeros(16, 32, device="cuda")
k[(1,)](d, out, 5, 7, 16, 32)
torch.cuda.synchronize()
print("triton TMA ok:", torch.equal(out, x[7:23, 5:37]))
