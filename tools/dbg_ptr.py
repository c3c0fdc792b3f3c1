import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from ssspy_b200 import _device
from ssspy_b200.bss import AuxLaplaceIVA
X = (np.random.randn(2,2,33,40)+1j*np.random.randn(2,2,33,40)).astype(np.complex64)
Xt = torch.from_numpy(X).cuda()
print("Xt", Xt.data_ptr(), Xt.is_contiguous(), Xt.dtype, Xt.is_cuda)
y = _device.to_device(Xt, torch.complex64); print("to_device", y.data_ptr(), y is Xt)
m = AuxLaplaceIVA(); m.input = Xt; print("after setter", m._dX.data_ptr(), m._dX is Xt)
m._reset(); print("after reset", m._dX.data_ptr())
