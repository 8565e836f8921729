import ctypes as C, numpy as np, torch
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tempme_b200 import _lib
L=_lib.lib()
for K,N in ((32,128),(32,64),(64,128),(128,128)):
    A=torch.randn(128,K,device='cuda'); B=torch.randn(N,K,device='cuda'); Cc=torch.empty(128,N,device='cuda')
    for _ in range(2):
        L.tm_selftest_gemm(_lib.ptr(A),_lib.ptr(B),_lib.ptr(Cc),K,N,2,None); torch.cuda.synchronize()
        L.tm_selftest_gemm(_lib.ptr(A),_lib.ptr(B),_lib.ptr(Cc),K,N,3,None); torch.cuda.synchronize()
