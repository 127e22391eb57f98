import sys, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from isaacgym_b200 import _native as N
from isaacgym_b200.config import CONFIGS
from isaacgym_b200.synth import make_state, clone_state
from oracle import task_oracle
for variant in ("tilt","a4","adof"):
    cfg = CONFIGS[variant]
    worst = 0.0; worst_rel=0.0
    for seed in (1,2,3):
        st = make_state(cfg, 8192, seed=seed)
        o = clone_state(st); task_oracle.post_physics_step(cfg, o)
        o64 = {k:(v.double() if v.dtype==torch.float32 else v.clone()) for k,v in st.items()}
        g = clone_state(st, "cuda"); g["stats"]=torch.zeros(64,8,dtype=torch.float64,device="cuda"); g["scratch"]=torch.zeros(16,dtype=torch.int32,device="cuda")
        N.check(N.load().ppk_post_physics_step(N.make_task(cfg), N.make_buffers(cfg,g), N.PHASE_ALL, None),"x"); torch.cuda.synchronize()
        a = g["obs_buf"].cpu().double().reshape(-1, cfg.num_obs); b = o["obs_buf"].double().reshape(-1, cfg.num_obs)
        J=len(cfg.body_ids); sl = slice(0, 6*J)
        scale = b.abs().amax(dim=-1, keepdim=True).clamp_min(1.0)
        err = ((a[:,sl]-b[:,sl]).abs()/scale).max().item()
        rel = ((a[:,sl]-b[:,sl]).abs()/(1e-5*b[:,sl].abs()+1e-6*scale)).max().item()
        worst=max(worst,err); worst_rel=max(worst_rel,rel)
    print(variant, "max |err|/rowscale on rotated fields: %.3e ; max err/tolerance: %.3f" % (worst, worst_rel))
