import torch, time
torch.backends.cudnn.benchmark = True
d = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
B = 8
x = torch.randn(B, 3, 1024, 1024, device=d).half()
pe = torch.nn.Conv2d(3, 1280, 16, 16).to(d).half()
c1 = torch.nn.Conv2d(1280, 256, 1, bias=False).to(d).half()
c3 = torch.nn.Conv2d(256, 256, 3, padding=1, bias=False).to(d).half()
tok = torch.randn(B, 64, 64, 1280, device=d).half()
y1 = torch.randn(B, 256, 64, 64, device=d).half()
with torch.no_grad():
    print("patch embed conv (NCHW)      us", t(lambda: pe(x)))
    print("patch embed + permute + add  us", t(lambda: pe(x).permute(0, 2, 3, 1) + tok))
    print("neck conv1x1 on permuted view us", t(lambda: c1(tok.permute(0, 3, 1, 2))))
    print("neck conv3x3 NCHW             us", t(lambda: c3(y1)))
    y1cl = y1.contiguous(memory_format=torch.channels_last)
    c3cl = c3.to(memory_format=torch.channels_last)
    print("neck conv3x3 channels_last    us", t(lambda: c3cl(y1cl)))
