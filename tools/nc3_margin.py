"""Margins of tests/test_gpu_train_step.py::test_train_step_with_a_class_count_that_is_not_a_multiple_of_8 in bf16: three runs of the 3-class
training step against the CPU oracle (relative loss error, cosine of the cv3 weight gradient).  usage (GPU box): python tools/nc3_margin.py"""
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch, numpy as np
from oracle import model as om, synth, cases
from yolo_ad_refine_b200.trainer import TrainEngine
nc=3
spec=[[k,([nc]+list(sh[1:]) if k in ("model.33.cv3.weight","model.33.cv3.bias") else sh),dt] for k,sh,dt in synth.load_spec()]
sd=synth.make_state_dict(seed=5,spec=spec)
img,bi,cl,bb=cases.train_step_inputs(**cases.TRAIN_STEP_CASES["b2_160"])
cl=(cl%nc).astype(cl.dtype)
t=[torch.from_numpy(a) for a in (img,bi,cl,bb)]
loss,items,grads,_,_=om.train_step_grads(sd,*t)
for rep in range(3):
    eng=TrainEngine(sd,dtype=torch.bfloat16,conv_impl=0,nc=nc)
    o=eng.forward_backward(t[0].cuda(),t[1],t[2],t[3]).cpu().numpy()
    g=eng.tp.g("model.33.cv3.weight").cpu(); r=grads["model.33.cv3.weight"]
    print(rep, abs(o[3]-float(loss))/float(loss), float(torch.nn.functional.cosine_similarity(g.reshape(-1),r.reshape(-1),dim=0)))
