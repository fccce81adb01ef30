import numpy as np
rng=np.random.default_rng(0)
kLboA=128*16+16; kKC=32; kKpAHalf=(kKC//4)*kLboA
def run(M,N,K,conjA,conjB):
    A=rng.standard_normal((M,K))+1j*rng.standard_normal((M,K))
    B=rng.standard_normal((K,N))+1j*rng.standard_normal((K,N))
    ns_tiles=(M+63)//64; per=(M+ns_tiles-1)//ns_tiles; N_t=((per+15)//16)*16
    ms_tiles=(2*N+127)//128; NKC=(K+15)//16
    b_half=N_t*kKC*4; stage_bytes=2*kKpAHalf+2*b_half
    sa=-1. if conjA else 1.; sb=-1. if conjB else 1.
    C=np.zeros((M,N),complex)
    for ms in range(ms_tiles):
      for ns in range(ns_tiles):
        D=np.zeros((128,N_t))
        for kc in range(NKC):
            st=np.full(stage_bytes//4,np.nan)  # floats (hi only: use full value at hi, lo=0)
            for ltid in range(256):
                wn=ltid&63; wkp0=ltid>>6
                w_so=wkp0*kLboA+(2*wn)*16
                n=ms*64+wn
                for j in range(2):
                    k=kc*16+2*(wkp0+4*j)
                    e0=B[k,n] if (n<N and k<K) else 0
                    e1=B[k+1,n] if (n<N and k+1<K) else 0
                    e=(np.real(e0),np.imag(e0),np.real(e1),np.imag(e1))
                    d=(w_so+(4*j)*kLboA)//4
                    st[d:d+4]=(e[0],-sb*e[1],e[2],-sb*e[3])
                    st[d+4:d+8]=(sb*e[1],e[0],sb*e[3],e[2])
                for j in range(2):
                    idx=ltid+256*j
                    if idx>=8*N_t: continue
                    kp=idx//N_t; xm=idx-kp*N_t
                    x_so=2*kKpAHalf+kp*(N_t*16)+xm*16
                    m=ns*N_t+xm; k=kc*16+2*kp
                    e0=A[m,k] if (m<M and k<K) else 0
                    e1=A[m,k+1] if (m<M and k+1<K) else 0
                    st[x_so//4:x_so//4+4]=(np.real(e0),sa*np.imag(e0),np.real(e1),sa*np.imag(e1))
            # UMMA semantic
            nks=4 if kc<NKC-1 else (2*K-(NKC-1)*32+7)//8
            for ks in range(nks):
                a_base=ks*2*kLboA; b_base=2*kKpAHalf+ks*2*(N_t*16)
                Am=np.zeros((128,8)); Bm=np.zeros((N_t,8))
                for kk in range(8):
                    for r in range(128):
                        Am[r,kk]=st[(a_base+(kk//4)*kLboA+(r//8)*128+(r%8)*16+(kk%4)*4)//4]
                    for r in range(N_t):
                        Bm[r,kk]=st[(b_base+(kk//4)*(N_t*16)+(r//8)*128+(r%8)*16+(kk%4)*4)//4]
                assert not np.isnan(Am).any() and not np.isnan(Bm).any()
                D+=Am@Bm.T
        # epilogue
        for q in range(4):
            vals=D[q*32:(q+1)*32]  # lane -> row
            m0=ns*N_t; mcols=min(N_t,M-m0)
            for c0 in range(0,mcols,16):
                for u in range(8):
                    for lane in range(32):
                        odd=lane&1
                        v=vals[lane,c0:c0+16]; vp=vals[lane^1,c0:c0+16]
                        mine=v[2*u+1] if odd else v[2*u]
                        recv=(vp[2*u] if (lane^1)&1 else vp[2*u+1])
                        n=ms*64+((q*32+lane)>>1); c=c0+2*u+odd
                        if n<N and c<mcols:
                            C[m0+c,n]=complex(recv,mine) if odd else complex(mine,recv)
    Ar=np.conj(A) if conjA else A; Br=np.conj(B) if conjB else B
    err=abs(C-Ar@Br).max()
    print(M,N,K,conjA,conjB,err); assert err<1e-9
run(5,9,7,0,0); run(33,70,18,1,0); run(130,64,16,0,1); run(20,8,35,1,1)
