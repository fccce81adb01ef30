"""Index-math check (numpy, CPU) of tc_kpipe.cuh's row-class mode: shifted B images, 16-byte aligned loads from the boundary at
or before each row start, row-interleaved tiles.  A restatement of the loader / image-builder / epilogue arithmetic; it does not
run the kernel."""
import numpy as np
rng=np.random.default_rng(0)
kLboA=128*16+16; kKC=32; kKpAHalf=(kKC//4)*kLboA
def run(R,K,N,lda,grid):
    mem=rng.standard_normal(R*lda+8)          # flat tensor (rows of pitch lda); 16-byte aligned base = index 0
    B=rng.standard_normal((K,N))
    N_t=((N+15)//16)*16; NKC=(K+3+31)//32; half=N_t*kKC
    img=np.zeros(4*NKC*2*half)
    for cls in range(4):
        sh=(cls*(lda%4))%4
        for k in range(K):
            c=(k+sh)//32; kk=(k+sh)%32
            for n in range(N):
                img[cls*NKC*2*half + c*2*half + (kk//4)*N_t*4+n*4+kk%4]=B[k,n]
    m_tiles=4*((R+511)//512)
    gx=min(grid&~3,m_tiles)
    C=np.full((R,N),np.nan)
    Kp=K+3
    for bx in range(gx):
        cls=bx&3; sh=(cls*lda)&3; k_end=K+sh
        for tile in range(bx,m_tiles,gx):
            assert tile%4==cls
            D=np.zeros((128,N_t))
            for kc in range(NKC):
                st=np.full((2*kKpAHalf+2*N_t*kKC*4)//4,np.nan)
                st[2*kKpAHalf//4:2*kKpAHalf//4+half]=img[(cls*NKC+kc)*2*half:(cls*NKC+kc)*2*half+half]
                for ltid in range(256):
                    kq=ltid&7; rbase=ltid>>3
                    row0=(tile>>2)*512+4*rbase+cls; k0=kc*32+kq*4
                    src=row0*lda-sh+k0
                    for i in range(4):
                        v=[0.]*4
                        if 128*i<R-row0:
                            q=src+i*128*lda
                            if k0>=sh and k0+4<=k_end:
                                assert q%4==0 and q>=0
                                v=list(mem[q:q+4])
                            else:
                                for e in range(4):
                                    if sh<=k0+e<k_end: v[e]=mem[q+e]
                        d=(kq*kLboA+rbase*16+i*512)//4
                        st[d:d+4]=v
                nks=4 if kc<NKC-1 else (Kp-(NKC-1)*32+7)//8
                for ks in range(nks):
                    a_base=ks*2*kLboA; b_base=2*kKpAHalf+ks*2*(N_t*16)
                    Am=np.zeros((128,8)); Bm=np.zeros((N_t,8))
                    for kk in range(8):
                        for r in range(128): Am[r,kk]=st[(a_base+(kk//4)*kLboA+r*16+(kk%4)*4)//4]
                        for r in range(N_t): Bm[r,kk]=st[(b_base+(kk//4)*(N_t*16)+r*16+(kk%4)*4)//4]
                    assert not np.isnan(Am).any() and not np.isnan(Bm).any()
                    D+=Am@Bm.T
            for q in range(4):
                for lane in range(32):
                    grow=(tile>>2)*512+4*(q*32+lane)+(tile&3)
                    if grow<R:
                        for c0 in range(0,N,16):
                            for u in range(8):
                                if c0+2*u<N: C[grow,c0+2*u:c0+2*u+2]=D[q*32+lane,c0+2*u:c0+2*u+2]
    A=np.stack([mem[r*lda:r*lda+K] for r in range(R)])
    err=abs(C-A@B).max(); print(R,K,N,lda,err); assert err<1e-9
run(600,70,36,70+1,8); run(1030,83,10,83,12); run(700,130,22,134,148); run(515,37,6,37,4)
