"""C4 (example/bubble.jl) against the finite-temperature Lindhard function (Maldague average of the T=0 formula of
example/bubble.jl:24-38) for different chain lengths / burn-in: quantifies the start bias of short chains."""
import sys, math, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import mcintegration_jl_b200 as mci
from cases import case
b=case("bubble"); para=b["para"]
exact=[-0.09716126865601389, -0.09510537068597597, -0.08858162927185188, -0.07605300600570582]
for solver, walkers, burn in (("vegas",1,0),("vegasmc",64,0),("vegasmc",512,0),("vegasmc",4096,0),("vegasmc",32768,0),("mcmc",512,0),("mcmc",4096,0),
                              ("vegasmc",4096,1000),("vegasmc",4096,4000),("vegasmc",32768,4000),("mcmc",4096,4000)):
    res=mci.integrate(b["f"], solver=solver, neval=1e8, niter=10, seed=99, measure="bin_by_discrete", walkers=walkers, chain_burnin=burn, **b["cfg"]())
    m=np.ravel(res.mean[0]); s=np.ravel(res.stdev[0])
    print(solver, walkers, "burnin", burn, "steps/chain", None if walkers is None else int(1e8/16/walkers), "dev_sigma", [round(float((a-e)/c),2) for a,e,c in zip(m,exact,s)], "sigma", ["%.1e"%x for x in s], flush=True)
