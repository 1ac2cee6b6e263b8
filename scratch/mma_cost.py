import sys
sys.path.insert(0, '.')
from scratch.ablate_tc import t
for (M,K,N) in [(2052096,128,128),(2052096,192,64),(2052096,64,64),(2052096,320,64)]:
    a = t(M,K,N,True,True,62); b = t(M,K,N,True,True,63); f = t(M,K,N,True,True,0)
    ntile = (M+127)//128; nk = (K+31)//32
    mmas = ntile*nk*12/148
    print((M,K,N), "full %.3f  only-MMA %.3f  nothing %.3f  -> MMA cost %.3f ms = %.0f clk per MMA (%d MMAs per CTA)" % (f, a, b, a-b, (a-b)*1e-3*1.965e9/mmas, mmas))
