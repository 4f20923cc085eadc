import torch
M=65536
for N,K in ((1728,576),(576,576),(2304,576),(576,2304),(2304,2304)):
    x=torch.randn(M,K,device='cuda',dtype=torch.float16); w=torch.randn(N,K,device='cuda',dtype=torch.float16)
    torch.matmul(x,w.t()); torch.cuda.synchronize()
