import sys
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops
dev = "cuda"
dd = torch.randn(2560, 256, device=dev).bfloat16(); mt = torch.randn(256, device=dev)
slots = torch.full((256, 192), -1, dtype=torch.int32, device=dev); slots[:, :10] = torch.arange(10, device=dev, dtype=torch.int32)
a0 = torch.randn(3, 256, device=dev); tc = torch.zeros(192, dtype=torch.int32, device=dev); a1 = torch.randn(192, 256, device=dev)
out = torch.empty(256 * 192, 256, device=dev, dtype=torch.bfloat16)
dz = torch.randn(256 * 192, 256, device=dev).bfloat16()
dmt = torch.zeros(256, device=dev); da0 = torch.zeros(3, 256, device=dev)
for _ in range(3):
    ops.decoder_assemble_fwd(dd, 10, mt, slots, 256, 192, add0=a0, tok_class=tc, add1=a1, out=out)
    ops.decoder_assemble_bwd(dz, slots, 256, 192, 10, dmask_token=dmt, dadd0=da0, tok_class=tc)
torch.cuda.synchronize()
