"""CPU emulation (numpy / pure Python) of the batched input-stage kernels driven by the REAL host planning
(DeviceTrainTransform._run_batched) and compared with tests/golden/input_stage.npz: checks the window / job-table logic
without a GPU.  python tests/tools/emul_input_stage.py  (about a minute)"""
import sys, importlib, ctypes as C, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
dt=importlib.import_module('synthetic-to-real-semantic-segmentation_b200.dataloders.device_transforms')
L=importlib.import_module('synthetic-to-real-semantic-segmentation_b200._lib')
from oracle import input_stage as OI
PB=22
def arr(ptr, n, dtype=np.uint8):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8 if dtype==np.uint8 else C.c_int32 if dtype==np.int32 else C.c_float)), shape=(n,))
def clip8(v): return np.uint8(min(255,max(0,v>>PB)))
def emu(name,*a):
    if name=="s2r_resize_bilinear_u8_multi":
        tab,n,mx,st=a
        jobs=(L.ResizeJob*n).from_address(tab)
        for j in jobs:
            bnd=arr(j.bounds, 1<<20, np.int32); 
            if j.axis==1:
                for line in range(j.lines):
                    for xo in range(j.on):
                        xx=j.o0+xo; xmin,cnt=int(bnd[2*xx]),int(bnd[2*xx+1]); k=arr(j.kk+4*xx*j.ksize, j.ksize, np.int32)
                        for c in range(j.C):
                            acc=1<<(PB-1)
                            for x in range(cnt):
                                sx=(j.W-1-(xmin+x)) if j.flip else xmin+x
                                acc+=int(arr(j.inp+line*j.in_pitch+sx*j.C+c,1)[0])*int(k[x])
                            arr(j.out+(line*j.on+xo)*j.C+c,1)[0]=clip8(acc)
            else:
                for yo in range(j.on):
                    yy=j.o0+yo; ymin,cnt=int(bnd[2*yy]),int(bnd[2*yy+1]); k=arr(j.kk+4*yy*j.ksize, j.ksize, np.int32)
                    for x in range(j.lines):
                        acc=1<<(PB-1)
                        for y in range(cnt):
                            acc+=int(arr(j.inp+(ymin-j.base+y)*j.in_pitch+x,1)[0])*int(k[y])
                        arr(j.out+yo*j.lines+x,1)[0]=clip8(acc)
    elif name=="s2r_resize_nearest_u8_multi":
        tab,n,mx,st=a
        for j in (L.NearestJob*n).from_address(tab):
            xt=arr(j.xtab,1<<16,np.int32); yt=arr(j.ytab,1<<16,np.int32)
            for y in range(j.OH):
                for x in range(j.OW):
                    sx,sy=int(xt[j.x0+x]),int(yt[j.y0+y]); v=0
                    if sx>=0 and sy>=0: v=arr(j.inp+sy*j.W+((j.W-1-sx) if j.flip else sx),1)[0]
                    arr(j.out+y*j.OW+x,1)[0]=v
    elif name=="s2r_input_stage_u8_multi":
        tab,n,mean,std,lut,fill,H,W,st=a
        lutv=arr(lut,256)
        tabn=OI.normalize_to_tensor(np.arange(256,dtype=np.uint8).reshape(256,1,1).repeat(3,2).reshape(256,1,3))  # [3][256][1]
        for j in (L.StageJob*n).from_address(tab):
            for y in range(H):
                for x in range(W):
                    sy,sx0=j.y1+y,j.x1+x; inside= sy<j.Hs and sx0<j.Ws; sx=(j.Ws-1-sx0) if j.flip else sx0
                    if j.img:
                        px=[0,0,0]
                        if inside: px=[int(v) for v in arr(j.img+(sy*j.Ws+sx)*3,3)]
                        o=np.ctypeslib.as_array(C.cast(j.out_img,C.POINTER(C.c_float)),shape=(3*H*W,))
                        for c in range(3): o[c*H*W+y*W+x]=tabn[c,px[c],0]
                    if j.label:
                        v=float(fill)
                        if inside: v=float(lutv[arr(j.label+sy*j.Ws+sx,1)[0]])
                        np.ctypeslib.as_array(C.cast(j.out_label,C.POINTER(C.c_float)),shape=(H*W,))[y*W+x]=v
    elif name=="s2r_gaussian_blur3_u8_multi":
        # blur_rows_kernel then blur_cols_kernel of csrc/input_stage.cu, statement by statement
        tab,n,H,W,st=a
        def tap(l,c,r,ww,fw): return ((c*ww+(l+r)*fw+(1<<23))&0xffffffff)>>24
        def three(p0,x,last,ww,fw):
            lo=lambda j: 0 if j<0 else j
            hi=lambda j: last if j>last else j
            p1=lambda j: tap(p0(lo(j-1)),p0(j),p0(hi(j+1)),ww,fw)
            p2=lambda j: tap(p1(lo(j-1)),p1(j),p1(hi(j+1)),ww,fw)
            return tap(p2(lo(x-1)),p2(x),p2(hi(x+1)),ww,fw)
        for j in (L.BlurJob*n).from_address(tab):
            tmp=arr(j.tmp,H*W*3); out=arr(j.out,H*W*3)
            for i in range(H*W*3):
                c,x,y=i%3,(i//3)%W,i//(3*W)
                sy=j.y1+y
                def px(xx):
                    sx0=j.x1+xx
                    if not (sy<j.Hs) or sx0>=j.Ws: return 0
                    return int(arr(j.img+sy*j.Ws*3+c+((j.Ws-1-sx0) if j.flip else sx0)*3,1)[0])
                tmp[i]=three(px,x,W-1,j.ww,j.fw)
            pitch=W*3
            for i in range(H*pitch):
                xb,y=i%pitch,i//pitch
                out[i]=three(lambda yy: int(tmp[xb+yy*pitch]),y,H-1,j.ww,j.fw)
    return 0
L.call=emu
dt.L.call=emu
fix=np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'golden', 'input_stage.npz'))
names=[str(k) for k in fix['cases']]
ok=True
for a,b in zip(names[0::2],names[1::2]):
    da,db=[int(v) for v in fix[a+'_draw']],[int(v) for v in fix[b+'_draw']]
    tr=dt.DeviceTrainTransform(1,da[2])
    src=torch.from_numpy(np.stack([fix[a+'_src'],fix[b+'_src']])); tgt=torch.from_numpy(np.stack([fix[a+'_tgt'],fix[b+'_tgt']])); lab=torch.from_numpy(np.stack([fix[a+'_lab'],fix[b+'_lab']]))
    N,H,W,_=src.shape; cs=da[2]
    out={'src_image':torch.zeros(N,3,cs,cs),'tgt_image':torch.zeros(N,3,cs,cs),'src_label':torch.zeros(N,cs,cs)}
    plan=[]
    for d in (da,db):
        ow,oh=dt._scale_size(W,H,d[1]); plan.append((bool(d[0]),ow,oh,d[3],d[4]))
    keep=tr._run_batched(src,tgt,lab,plan,out,None)
    for n,k in enumerate((a,b)):
        e=[np.array_equal(out['src_image'][n].numpy(),fix[k+'_out_src']),np.array_equal(out['tgt_image'][n].numpy(),fix[k+'_out_tgt']),np.array_equal(out['src_label'][n].numpy(),fix[k+'_out_lab'])]
        print(k,plan[n],e); ok&=all(e)
# RandomGaussianBlur cases (the blur fires on both samples of a pair, own radius per image) and a mixed batch
names=[str(k) for k in fix['blur_cases']]
for a,b in zip(names[0::2],names[1::2]):
    da,db=[int(v) for v in fix[a+'_draw']],[int(v) for v in fix[b+'_draw']]
    tr=dt.DeviceTrainTransform(1,da[2])
    src=torch.from_numpy(np.stack([fix[a+'_src'],fix[b+'_src']])); tgt=torch.from_numpy(np.stack([fix[a+'_tgt'],fix[b+'_tgt']])); lab=torch.from_numpy(np.stack([fix[a+'_lab'],fix[b+'_lab']]))
    N,H,W,_=src.shape; cs=da[2]
    out={'src_image':torch.zeros(N,3,cs,cs),'tgt_image':torch.zeros(N,3,cs,cs),'src_label':torch.zeros(N,cs,cs)}
    plan,blur=[],[]
    for d,k in ((da,a),(db,b)):
        ow,oh=dt._scale_size(W,H,d[1]); plan.append((bool(d[0]),ow,oh,d[3],d[4]))
        r=fix[k+'_radii']
        blur.append({'src_image':dt._gaussian_blur_weights(float(r[0])),'tgt_image':dt._gaussian_blur_weights(float(r[1]))})
    keep=tr._run_batched(src,tgt,lab,plan,out,None,blur)
    for n,k in enumerate((a,b)):
        e=[np.array_equal(out['src_image'][n].numpy(),fix[k+'_out_src']),np.array_equal(out['tgt_image'][n].numpy(),fix[k+'_out_tgt']),np.array_equal(out['src_label'][n].numpy(),fix[k+'_out_lab'])]
        print(k,plan[n],e); ok&=all(e)
print("ALL OK" if ok else "MISMATCH")
