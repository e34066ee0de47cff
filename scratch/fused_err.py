import numpy as np, torch, sys
sys.path.insert(0,'tests'); sys.path.insert(0,'oracle'); sys.path.insert(0,'.')
import plan_emulator as PE, talkshow_oracle as O
from talkshow_b200 import _lib, synth
from talkshow_b200.engine import Engine
pix = synth.body_pixel_checkpoint(seed=0)
sd = pix["generator"]
B,T=3,8
g=torch.Generator().manual_seed(3)
codes=torch.randint(0,2048,(B,T,2),generator=g); label=torch.tensor([0,3,1])
aud=O.audio_encoder(pix["audioencoder"], synth.synth_mfcc(B,4*T,seed=5))
ref=O.pixelcnn_forward(sd,codes,label,aud.unsqueeze(-1).repeat(1,1,1,2)).numpy()
a = torch.einsum("oc,bct->bto", sd["embedding_aud.weight"][:, :, 0, 0], aud) + sd["embedding_aud.bias"]
av = (torch.einsum("oc,btc->bto", sd["fusion_v.weight"][:, 256:, 0, 0], a) + sd["fusion_v.bias"]).numpy()
ah = (torch.einsum("oc,btc->bto", sd["fusion_h.weight"][:, 256:, 0, 0], a) + sd["fusion_h.bias"]).numpy()
res={}
for fused in (1,0):
    e=Engine(-148); e.set_pixelcnn_fusion(fused); e.load_pixelcnn(sd)
    table,blob=_lib.plan_to_numpy(e.h); print(fused, 'staged', e.L.ts_pixelcnn_staged_row_bytes(e.h)); e.close()
    p=PE.Plan(table,blob)
    cls_w=[sd["layers.%d.class_cond_embedding.weight"%l].numpy() for l in range(p.L)]
    _,lg=PE.run(p,sd["embedding.weight"].numpy(),cls_w,av,ah,label.numpy(),codes.numpy(),T)
    got=np.transpose(lg,(2,3,0,1)); res[fused]=got
    print('fused' if fused else 'plain', p.nstages, 'max err', np.abs(got-ref).max(), 'logit std', ref.std())
print('fused vs plain', np.abs(res[1]-res[0]).max())
