"""Throughput of the reference options other than the plain spectrogram (waterfall / turnFlip, split-real / channelMode) against
the plain render, device resident.  usage (under gpurun): python tools/mode_bench.py"""
import sys, os, json
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spectroplot-js_b200")): sys.path.insert(0, p)
import numpy as np, torch, spectro_b200
from spectro_b200 import windows, cmaps
dev=torch.device("cuda",0); eng=spectro_b200.Engine(0)
st=torch.cuda.Stream(device=dev); torch.cuda.set_stream(st); eng.set_stream(st.cuda_stream)
cm=cmaps.cmap_bytes([list(c) for c in cmaps.cmaps["viridis_cmap"]])
for fmt,n,S,wf,chm in (("CS16",4096,1<<26,False,False),("CS16",4096,1<<26,True,False),("CS16",4096,1<<26,False,True),("CU8",1024,1<<26,True,False),("CU8",1024,1<<26,False,False),("CU8",1024,1<<26,False,True),("CS16",512,1<<26,False,True),("CS16",128,1<<26,False,True)):
    sw=4 if fmt=="CS16" else 2; width=S//n
    d_in=torch.empty(S*sw+256,dtype=torch.uint8,device=dev); eng.synth_fill(d_in.data_ptr(),fmt,0,S,S,5)
    d_img=torch.empty(4*width*n,dtype=torch.uint8,device=dev); d_g=torch.empty(3*width,dtype=torch.uint8,device=dev)
    d_h=torch.zeros(1256,dtype=torch.int64,device=dev); d_mm=torch.zeros(2,dtype=torch.float64,device=dev)
    w=windows.hannWindow(n); ww=np.array(w["window"])
    def step():
        rq,keep=eng.make_request(d_in.data_ptr(),fmt,n,width,ww,1/w["weight"],6,30,cm,channel_mode=chm,waterfall=wf,byte_length=S*sw)
        return eng.render_enqueue(rq,d_img.data_ptr(),(d_g.data_ptr(),d_g.data_ptr()+width,d_g.data_ptr()+2*width),d_h.data_ptr(),d_h.data_ptr()+8000,d_mm.data_ptr())
    for _ in range(3): rp=step()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5): rp=step()
    e1.record(st); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/5; eng.render_finish(rp)
    print(fmt,n,"waterfall" if wf else "spectrogram","split" if chm else "", round(ms,3),"ms", round(S/ms/1e6,1),"GS/s", flush=True)
