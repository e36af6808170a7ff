import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from memflow_b200.phasespace.phasespace import PhaseSpace
E=13000.0; M=[125.25,172.5,172.5]
ps=PhaseSpace(E,[21,21],[25,6,-6],final_masses=torch.tensor(M,dtype=torch.float64),dev="cuda:0")
u=torch.full((64,7),0.5,dtype=torch.float64); u[1,0]=0.0; u[2,0]=1.0; u[3,5]=1.0
mom,w,x1,x2=ps.get_momenta_from_ps(u.cuda())
om,ow,ox1,ox2=O.rambo_forward(u.numpy(),M,E)
d=np.abs(mom.cpu().numpy()-om).reshape(64,-1)
print("rows with diff:", [(i, float(np.nanmax(d[i]))) for i in range(6)])
print("w", w[:5].cpu().numpy(), ow[:5])
print(mom[1].cpu().numpy()); print(om[1])
# get_ps
from memflow_b200.unfolding_flow.utils import Compute_ParticlesTensor
g=np.load("tests/golden/get_ps_n4.npz")
mom_,boost=torch.from_numpy(g["mom"]).cuda(),torch.from_numpy(g["boost"]).cuda()
psx,jac,mask=Compute_ParticlesTensor.get_PS(mom_,boost,order=[0,1,2,3],target_mass=torch.from_numpy(g["target_mass"]),E_CM=E)
print("mask eq", torch.equal(mask.cpu(), torch.from_numpy(g["mask_id"])), int(mask.sum()), int(g["mask_id"].sum()))
dd=np.abs(psx.cpu().numpy()-g["ps_id"]); good=~g["mask_id"]
print("ps diff good max", np.nanmax(dd[good]), "q99", np.nanquantile(dd[good],0.99), "argmax", np.unravel_index(np.nanargmax(np.where(good[:,None],dd,0)),dd.shape))
print("nan eq", np.array_equal(np.isnan(psx.cpu().numpy()),np.isnan(g["ps_id"])), "jac nan eq", np.array_equal(np.isnan(jac.cpu().numpy()),np.isnan(g["jac_id"])))
