import numpy as np
rng=np.random.default_rng(0)
kLboA=128*16+16; kKC=32; kKpAHalf=(kKC//4)*kLboA
def run(O,H,J,I):
    X=rng.standard_normal((O,H,I))+1j*rng.standard_normal((O,H,I))
    Mat=rng.standard_normal((J,H))+1j*rng.standard_normal((J,H))
    N=2*J;K=2*H
    n_tiles=(N+255)//256; per=(N+n_tiles-1)//n_tiles; N_t=((per+15)//16)*16
    NKC=(K+31)//32; half=N_t*kKC
    img=np.zeros(n_tiles*NKC*2*half)
    for j in range(J):
        for hh in range(H):
            re,im=Mat[j,hh].real,Mat[j,hh].imag
            for ci in range(2):
                for co in range(2):
                    b=re if ci==co else (-im if ci==1 else im)
                    k=2*hh+ci;n=2*j+co;t=n//N_t;nl=n%N_t;c=k//32;kk=k%32
                    base=(t*NKC+c)*2*half
                    img[base+(kk//4)*N_t*4+nl*4+kk%4]=b
    R=O*I; m_tiles=(R+127)//128
    b_half=N_t*kKC*4; stage_bytes=2*kKpAHalf+2*b_half
    Y=np.zeros((O,J,I),complex)
    Xf=X.reshape(-1)
    for nt in range(n_tiles):
      for tile in range(m_tiles):
        D=np.zeros((128,N_t))
        for kc in range(NKC):
            st=np.full(stage_bytes//4,np.nan)
            st[2*kKpAHalf//4:2*kKpAHalf//4+half]=img[(nt*NKC+kc)*2*half:(nt*NKC+kc)*2*half+half]
            for ltid in range(256):
                rl=ltid&127;hh=ltid>>7
                so=hh*kLboA+rl*16
                r=tile*128+rl
                for j in range(4):
                    h=kc*16+2*hh+4*j
                    e0=e1=0
                    if r<R:
                        o=r//I;i=r-o*I
                        base=(o*H*I+i)   # complex index
                        if h<H: e0=Xf[base+h*I]
                        if h+1<H: e1=Xf[base+(h+1)*I]
                    d=(so+(2*j)*kLboA)//4
                    st[d:d+4]=(np.real(e0),np.imag(e0),np.real(e1),np.imag(e1))
            nks=4 if kc<NKC-1 else (K-(NKC-1)*32+7)//8
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
        jt0=nt*(N_t//2); jn=min(N_t//2,J-jt0)
        for q in range(4):
            for lane in range(32):
                r=tile*128+q*32+lane
                if r>=R: continue
                o=r//I;i=r-o*I
                for c0 in range(0,2*jn,32):
                    second=c0+16<N_t
                    for t in range(16):
                        j=c0//2+t
                        if j<jn and (t<8 or second):
                            col=c0+(t>>3)*16+2*(t&7)
                            Y[o,jt0+j,i]=complex(D[q*32+lane,col],D[q*32+lane,col+1])
    ref=np.einsum('jh,ohi->oji',Mat,X)
    err=abs(Y-ref).max(); print(O,H,J,I,N_t,n_tiles,err); assert err<1e-9
run(3,5,7,11); run(10,36,150,18); run(8,33,12,20); run(2,17,129,70)
