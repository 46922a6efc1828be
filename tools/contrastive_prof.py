import sys, ctypes as C
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sfv_b200
def down3(n):
    for _ in range(3): n=(n-1)//2+1
    return n
R,L=512,25
for prec in ("mixed","bf16","fp32"):
    rb = sfv_b200.Seq2SeqBinaryVAE(3,3,L,L,kind="contrastive",input_hw=(R,R),precision=prec)
    rb.load_state_dict(sfv_b200.init_rbvae_state_dict(3,L,(down3(R),down3(R)),channels=64,num_layers=2,seed=2))
    x=torch.rand(64,1,3,R,R,device="cuda")
    for _ in range(3): rb.encode_codes(x, noise_ratio=0.0)
    torch.cuda.synchronize()
    lib=sfv_b200.lib(); lib.sfv_profile_enable(1)
    rb.encode_codes(x, noise_ratio=0.0); torch.cuda.synchronize()
    for cat,name in enumerate(["tc_gemm","conv_in/igemm","gn_stats","gn_apply","softmax","other"]):
        ms,work,n=C.c_double(),C.c_double(),C.c_int64()
        lib.sfv_profile_read(cat,C.byref(ms),C.byref(work),C.byref(n))
        if n.value: print(prec,name,round(ms.value,3),"ms",n.value,"launches")
    print(lib.sfv_profile_log().decode())
    lib.sfv_profile_enable(0)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): rb.encode_codes(x, noise_ratio=0.0)
    e1.record(); torch.cuda.synchronize(); print(prec,"total per call",e0.elapsed_time(e1)/5)
