"""Shared-memory bank conflicts of the rotate gather for candidate lane -> pixel mappings and tile pitches
(wavefronts per LDS.U8 / STS.32 instruction; model: max distinct 32-bit words per bank per warp instruction)."""
import numpy as np, math
S=128
def desc(crop, theta_deg):
    angle=np.float32(theta_deg)
    rad=float(angle)*0.01745329251994329
    s,c=math.sin(rad),math.cos(rad)
    w=h=float(crop)
    cxx,cyy,sx,sy=c*w,c*h,s*w,s*h
    nx=int(max(abs(cxx+sy),abs(cxx-sy),abs(-cxx+sy),abs(-cxx-sy)))
    ny=int(max(abs(sx+cyy),abs(sx-cyy),abs(-sx+cyy),abs(-sx-cyy)))
    isin=int(s*65536.0); icos=int(c*65536.0)
    ax=(nx<<15)-int(c*float((nx-1)<<15)); ay=(ny<<15)-int(s*float((nx-1)<<15))
    xd=(crop-nx)*32768; yd=(crop-ny)*32768
    return nx,ny,isin,icos,ax+xd,ay+yd,ny//2
def wavefronts(crop,pitch,theta,patch_w_lanes,anchor=(64,64),fov_pitch=128):
    nx,ny,isin,icos,rax,ray,rcy=desc(crop,theta)
    left=anchor[0]-(nx>>1); top=anchor[1]-(ny>>1)
    lw=patch_w_lanes; lh=32//lw
    tot=0; n=0; st=0
    for py in range(0,S,lh):
        for px in range(0,S,4*lw):
            lanes=[(px+4*(t%lw), py+t//lw) for t in range(32)]
            for k in range(4):
                banks={}
                for (ox,oy) in lanes:
                    rxp=ox+k-left; ryp=oy-top
                    dx=rax+isin*(rcy-ryp)+rxp*icos; dy=ray-icos*(rcy-ryp)+rxp*isin
                    a=(dy>>16)*pitch+(dx>>16)
                    wd=a>>2
                    banks.setdefault(wd&31,set()).add(wd)
                tot+=max(len(v) for v in banks.values()); n+=1
            banks={}
            for (ox,oy) in lanes:
                wd=(oy*fov_pitch+ox)>>2
                banks.setdefault(wd&31,set()).add(wd)
            st+=max(len(v) for v in banks.values())
    return tot/n, st/(n/4)
for crop,pitches in ((182,(208,)),(230,(256,272))):
    for pitch in pitches:
        for lw in (32,8,4,2,1):
            for fp in (128,144):
                r=[wavefronts(crop,pitch,th,lw,fov_pitch=fp) for th in (3,17,33,45,61,80,90.5,100,135,170)]
                print(f"crop {crop} pitch {pitch} lanes_w {lw:2d} fov_pitch {fp}: LDS avg {np.mean([a for a,b in r]):.2f} max {max(a for a,b in r):.2f} | STS {np.mean([b for a,b in r]):.2f}")
