import torch, sys
sys.path.insert(0,'/root/repo')
dev=torch.device('cuda:0')
n=8192*300*8
xh=torch.empty(n,dtype=torch.float32).pin_memory(); oh=torch.empty(n,dtype=torch.float32).pin_memory()
d1=torch.empty(n,device=dev); d2=torch.empty(n,device=dev)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(fn,k=10):
    fn(); torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/k
def both(nch):
    cur=torch.cuda.current_stream(); s1.wait_stream(cur); s2.wait_stream(cur)
    c=n//nch
    for i in range(nch):
        sl=slice(i*c,(i+1)*c)
        with torch.cuda.stream(s1): d1[sl].copy_(xh[sl],non_blocking=True)
        with torch.cuda.stream(s2): oh[sl].copy_(d2[sl],non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
for nch in (1,2,4,8,16,32,64):
    print(f'{nch} chunks, H2D and D2H concurrently, no kernels: {t(lambda: both(nch)):.3f} ms', flush=True)
# with a bandwidth-hungry kernel running concurrently (device copy loop)
big=torch.empty(256*1024*1024,device=dev,dtype=torch.float32); big2=torch.empty_like(big)
s3=torch.cuda.Stream()
def both_with_traffic(nch):
    cur=torch.cuda.current_stream(); s3.wait_stream(cur)
    with torch.cuda.stream(s3):
        for _ in range(4): big2.copy_(big)
    both(nch); cur.wait_stream(s3)
print(f'16 chunks with concurrent device-to-device copies (HBM busy): total {t(lambda: both_with_traffic(16),k=5):.3f} ms (4 x 2 GB d2d alone: {t(lambda: [big2.copy_(big) for _ in range(4)],k=5):.3f} ms)')
# pipeline structure with a trivial kernel in place of the LM step: H2D chunk -> event -> kernel -> event -> D2H chunk
s_in,s_out=torch.cuda.Stream(),torch.cuda.Stream(); s_run=[torch.cuda.Stream() for _ in range(4)]
def piped(nch, work):
    cur=torch.cuda.current_stream()
    for s in [s_in,s_out]+s_run: s.wait_stream(cur)
    c=n//nch
    for i in range(nch):
        sl=slice(i*c,(i+1)*c)
        e1=torch.cuda.Event(); e2=torch.cuda.Event()
        with torch.cuda.stream(s_in): d1[sl].copy_(xh[sl],non_blocking=True); e1.record(s_in)
        sr=s_run[i%4]
        with torch.cuda.stream(sr):
            sr.wait_event(e1); work(d1[sl],d2[sl]); e2.record(sr)
        with torch.cuda.stream(s_out):
            s_out.wait_event(e2); oh[sl].copy_(d2[sl],non_blocking=True)
    for s in [s_in,s_out]+s_run: cur.wait_stream(s)
print(f'pipeline structure, 16 chunks, trivial kernel (d2 = d1 * 2): {t(lambda: piped(16, lambda a,b: torch.mul(a,2,out=b))):.3f} ms')
def heavy(a,b):
    for _ in range(30): torch.mul(a,2,out=b)
print(f'pipeline structure, 16 chunks, 30 elementwise kernels per chunk: {t(lambda: piped(16, heavy)):.3f} ms')
def sleepy(a,b):
    torch.cuda._sleep(int(0.28e-3*1.9e9)); torch.mul(a,2,out=b)
print(f'pipeline structure, 16 chunks, 0.28 ms spin kernel per chunk (latency like the LM step, no memory traffic): {t(lambda: piped(16, sleepy)):.3f} ms')
# two copy streams per direction, alternating chunks (can the copy engine's wait for chunk c+1 overlap the copy of chunk c?)
s_in2=[torch.cuda.Stream() for _ in range(2)]; s_out2=[torch.cuda.Stream() for _ in range(2)]
def piped2(nch, work, n_in=2, n_out=2):
    cur=torch.cuda.current_stream()
    for s in s_in2+s_out2+s_run: s.wait_stream(cur)
    c=n//nch
    for i in range(nch):
        sl=slice(i*c,(i+1)*c)
        e1=torch.cuda.Event(); e2=torch.cuda.Event()
        si=s_in2[i%n_in]; so=s_out2[i%n_out]
        with torch.cuda.stream(si): d1[sl].copy_(xh[sl],non_blocking=True); e1.record(si)
        sr=s_run[i%4]
        with torch.cuda.stream(sr):
            sr.wait_event(e1); work(d1[sl],d2[sl]); e2.record(sr)
        with torch.cuda.stream(so):
            so.wait_event(e2); oh[sl].copy_(d2[sl],non_blocking=True)
    for s in s_in2+s_out2+s_run: cur.wait_stream(s)
triv=lambda a,b: torch.mul(a,2,out=b)
for n_in,n_out in ((1,1),(1,2),(2,1),(2,2)):
    print(f'pipeline structure, 16 chunks, trivial kernel, {n_in} copy-in / {n_out} copy-out streams: {t(lambda: piped2(16, triv, n_in, n_out)):.3f} ms')
