import sys, copy, torch, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from test_ppo_trunk_gpu import _nets, _cpu, _rel
cuda=torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32=True
import os
for L,N in ((10,77),(8,300),(4,128)):
    for v1 in (0,1):
        if v1: os.environ['MM_PPO_TRUNK_V1']='1'
        else: os.environ.pop('MM_PPO_TRUNK_V1',None)
        policy=_nets(cuda,seed=L)
        g=torch.Generator().manual_seed(100+N)
        obs=torch.rand((L*N,47),generator=g); h0=torch.randn((N,1,32),generator=g)*0.5
        masks=(torch.rand((L*N,1),generator=g)>0.2).float()
        act=torch.randint(0,5,(L*N,1),generator=g).float()
        wl,wv=torch.randn((L*N,1),generator=g),torch.randn((L*N,1),generator=g)
        dev=lambda x:x.to(cuda)
        for name,net in (('actor',policy.actor),('critic',policy.critic)):
            ref=_cpu(net)
            if name=='actor':
                lp_r,ent_r=ref.evaluate_actions(obs,h0,act,masks,active_masks=torch.ones(L*N,1)); (lp_r*wl).sum().add(ent_r*3.0).backward()
                lp_f,ent_f=net.evaluate_actions(dev(obs),dev(h0),dev(act),dev(masks),active_masks=torch.ones(L*N,1,device=cuda)); (lp_f*dev(wl)).sum().add(ent_f*3.0).backward()
            else:
                v_r,_=ref(obs,h0,masks); (v_r*wv).sum().backward()
                v_f,_=net(dev(obs),dev(h0),dev(masks)); (v_f*dev(wv)).sum().backward()
            errs={k:round(_rel(pf.grad.cpu(),pr.grad),4) for (k,pr),pf in zip(ref.named_parameters(),net.parameters()) if pr.grad is not None}
            print('v1' if v1 else 'v2',L,N,name,errs,flush=True)
