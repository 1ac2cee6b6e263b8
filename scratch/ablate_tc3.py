import sys
sys.path.insert(0, '.')
from scratch.ablate_tc import t
names = [(0,"full"),(1,"noMMA"),(2,"noLDG"),(4,"noSTS"),(8,"noEPIst"),(16,"noEPIbody"),(32,"noLoaderMath"),(1|16,"noMMA+noEPIbody"),(2|4|32,"no loader work"),(63,"nothing")]
for (M,K,N) in [(2052096,192,64),(2052096,320,64),(2052096,128,128)]:
    print("shape", (M,K,N), "ideal HBM ms %.3f" % (4*(M*K+M*N)/6.5237e9))
    for f,nm in names:
        print("   %-18s %.3f ms" % (nm, t(M,K,N,True,True,f)), flush=True)
