import sys, os, torch
sys.path.insert(0,'benchmarks'); sys.path.insert(0,'pagedattention-based-transformer-decoder-inference-framework_b200')
import llm_decoder as ld
dev=torch.device('cuda',0); torch.cuda.set_device(dev)
M=int(os.environ.get("M","2048")); HID,INTER=4096,16384
g=torch.Generator(device=dev).manual_seed(1)
W1=[torch.randint(-127,128,(1,HID,INTER),generator=g,device=dev,dtype=torch.int8) for _ in range(2)]
x=torch.randint(-127,128,(1,M,HID),generator=g,device=dev,dtype=torch.int8)
y1=torch.empty((1,M,INTER),dtype=torch.int8,device=dev); b1=torch.randn(INTER,device=dev)
for i in range(4):
    assert ld.dnnl_matmul_int8(x,W1[i%2],y1,1,M,INTER,HID,1/16,1/16,8.0,b1,"relu")
torch.cuda.synchronize(); print("done")
