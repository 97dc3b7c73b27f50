import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from tests.common import *
from luci_b200 import engine
g=load_golden('fit_sincgauss_groups')
cfg=config_from_golden(g)
dev=torch.device('cuda',0)
t=lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
res=engine.fit_batch(cfg,t(g['cube'].astype(np.float32)),t(g['theta'].astype(np.float32)),t(g['v0'].astype(np.float32)),t(g['b0'].astype(np.float32)),want_sol=True)
rows=res['rows'].cpu().numpy(); st=res['status'].cpu().numpy()
rep=parity_report(rows,g,g['model'])
L=5
print('vel_rel',g['vel_rel'],'sigma_rel',g['sigma_rel'])
for i in range(len(st)):
    print(i,'it',st[i]&255,'ev',(st[i]>>8)&255,'flags',hex(st[i]>>16),'dchi %.3e'%rep['dchi'][i],'dv %.3f'%rep['dv'][i],'df %.2e'%rep['df'][i])
for i in (1,2):
    print('spaxel',i)
    print(' ours vel',rows[i,3*L:4*L],'\n ref  vel',g['ref_velocities'][i])
    print(' ours sig',rows[i,5*L:6*L],'\n ref  sig',g['ref_sigmas'][i])
    print(' ours amp',rows[i,0:L],'\n ref  amp',g['ref_amplitudes'][i])
    print(' chi2',rows[i,7*L],g['ref_chi2'][i], 'v0',g['v0'][i],'b0',g['b0'][i])
