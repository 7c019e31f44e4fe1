import sys; import pathlib; R_=pathlib.Path(__file__).resolve().parents[2]; sys.path.insert(0,str(R_/'tests')); sys.path.insert(0,str(R_))
import importlib, numpy as np, oracle_lib as O, emu_lib
from conftest import oracle_bias
from test_emu_parity import assert_walk_matches
pkg=importlib.import_module("voxel-raycaster_b200"); S=pkg.scene
rng=np.random.default_rng(int(sys.argv[1])); N=int(sys.argv[2]); bad=0
for it in range(N):
    n=int(rng.choice([128,256]))
    vol=np.zeros((n,n,n),np.int8)
    dens=rng.choice([0.0002,0.001,0.005])
    m=rng.random((n,n,n))<dens; vol[m]=5
    vol[rng.random((n,n,n))<dens*0.2]=6
    if rng.random()<0.5: vol[:n//8]=5       # a ground slab
    pos=(rng.random(3)*n).astype(np.float32); pos[2]=max(pos[2], n//8+1.3)
    d=np.array([rng.random()*np.pi, rng.random()*2*np.pi],np.float32)
    nl=int(rng.choice([1,2]))
    lights=np.zeros((8,10),np.float32)
    for l in range(nl): lights[l]=[rng.random(),rng.random(),rng.random(),1.0,*(rng.random(3)*n),-1,-1,-1.5]
    sc=S.Scene(n,vol,96,64,pos,d,lights,max_distance=3*n)
    table=O.make_ray_table(96,64); desc,root=pkg.octree_generate(vol)
    ref_rgba,ref_aux,_=O.raycast(sc,table,octree=(desc,root),shadow_lights=nl)
    bias=oracle_bias(O,sc,desc,root)
    for svo in (1,2):
        rgba,aux=emu_lib.raycast(sc,table,bias=bias,use_svo=svo,shadow_lights=nl)
        try: assert_walk_matches(ref_rgba,ref_aux,rgba,aux,svo==2,f"it {it} svo {svo}")
        except AssertionError as e: bad+=1; print("MISMATCH",it,svo,n,dens,pos,d,nl,str(e)[:100])
    if it%20==0: print(it, "max steps", ref_aux["steps_total"].max(), "ties", ((ref_aux["flags"]&4)!=0).mean(), flush=True)
print("done",N,"bad",bad)
