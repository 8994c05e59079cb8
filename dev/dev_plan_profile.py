"""Dev tool (GPU): profile of TransformPlan construction at config-2 size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_b200 import plan as P
from scri_inputs import real_supertranslation
kw = dict(supertranslation=real_supertranslation(4), frame_rotation=[1.0, 2.0, 3.0, 4.0], boost_velocity=[0.01, 0.02, 0.03])
for _ in range(5): P.TransformPlan(2, 8, sb.h, r_is_scaled_out=True, **kw)
torch.cuda.synchronize()
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20): P.TransformPlan(2, 8, sb.h, r_is_scaled_out=True, **kw)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
