import os, sys, json, ctypes
sys.path.insert(0, '/root/repo'); os.environ["VSB_NO_APP_CONFIG_WRITE"]="1"
sys.path.insert(0, '/root/repo/voxel-stack-blender_b200')
import numpy as np, torch
import bench, vsb_engine as E, vsb_native as N, vsb_params as P, vsb_synth as S, lut_manager
W,H,L=15120,6230,12
cfg=bench.workload_config(bench.write_lut_file())
Z0=int(os.environ.get('Z0','20')); stack=S.DeviceStack(W,H,L,seed=1,z0=Z0,total_layers=max(436,Z0+L))
ops=P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params))
eng=E.StackEngine(W,H,cfg,ops)
outs=torch.empty((2,H,stack.pitch),dtype=torch.uint8,device='cuda')
eng.process_device_batch(stack.ptr(0),stack.pitch,stack.layer_stride,outs.data_ptr(),stack.pitch,H*stack.pitch,2,5)
N.load().vsb_layer_stats(1,None)
n=6
eng.process_device_batch(stack.ptr(5),stack.pitch,stack.layer_stride,outs.data_ptr(),stack.pitch,H*stack.pitch,2,n)
buf=(ctypes.c_ulonglong*16)(); N.load().vsb_layer_stats(2,buf)
v=[x/n for x in buf]
print("per layer: uniform tiles %.0f, worked tiles %.0f, rec tiles %.0f, rec px %.0f (%.2f%%), bilateral px %.0f (%.2f%%), gauss words %.0f (px %.2f%%)"%(v[0],v[1],v[2],v[3],100*v[3]/(W*H),v[4],100*v[4]/(W*H),v[5],400*v[5]/(W*H)))
ph=["prologue","P2 setup","search","P3+P4 planes","P5 tests","bilateral","gaussian"]
tot=sum(v[8:15]) or 1
print("thread-0 cycles per phase (share): "+", ".join("%s %.1f%%"%(n,100*c/tot) for n,c in zip(ph,v[8:15])), " total Mcycles/layer %.1f"%(tot/1e6))
g=stack.layers[6][:, :W]; print("white frac", float((g>127).float().mean()))
