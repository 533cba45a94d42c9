import sys, torch
sys.path.insert(0,'nerf-workspaces-explorer_b200'); sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from conftest import load_golden
from nwx import engine as E
g = load_golden("raw2outputs")
raw, z, d = (g[k].cuda() for k in ("raw","z_vals","rays_d"))
flags = torch.zeros(1, dtype=torch.int32, device='cuda')
out = E.composite(raw, z, d, None, False, flags=flags)
torch.cuda.synchronize()
print('flags', flags, 'nan disp', torch.isnan(out[1]).sum().item(), out[1][:6])
flags2 = torch.zeros(4, dtype=torch.int32, device='cuda')
out = E.composite(raw[:2].contiguous(), z[:2].contiguous(), d[:2].contiguous(), None, False, flags=flags2)
torch.cuda.synchronize(); print('flags2', flags2, out[1])
