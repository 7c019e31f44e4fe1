import sys; import pathlib; R_=pathlib.Path(__file__).resolve().parents[2]; sys.path.insert(0,str(R_/'tests')); sys.path.insert(0,str(R_))
import importlib, numpy as np, oracle_lib as O, emu_lib
from conftest import oracle_bias
from test_emu_parity import assert_walk_matches
pkg=importlib.import_module("voxel-raycaster_b200"); S=pkg.scene
bad=0
for n,variant in ((256,"shell"),(256,"solid"),(128,"shell"),(512,"shell")):
    vol=S.terrain_map(n,variant,reflect_fraction=0.02); h=S.heightfield(n)
    desc,root=pkg.octree_generate(vol); table=O.make_ray_table(192,108)
    for cam in range(0,40 if n < 512 else 16):
        pos,d=S.make_camera(n,h,cam,need_zero_bias=(cam%2==0))
        if cam%5==0: pos=pos+np.array([0,0,n*0.3],np.float32)     # high above the terrain: big empty cells, biased start
        for nl in (1,2):
            sc=S.Scene(n,vol,192,108,pos.astype(np.float32),d,S.make_lights(n,nl),max_distance=3*n)
            ref_rgba,ref_aux,_=O.raycast(sc,table,octree=(desc,root),shadow_lights=nl)
            bias=oracle_bias(O,sc,desc,root)
            for svo in (1,2):
                rgba,aux=emu_lib.raycast(sc,table,bias=bias,use_svo=svo,shadow_lights=nl)
                try: assert_walk_matches(ref_rgba,ref_aux,rgba,aux,svo==2,f"{n} {variant} cam {cam} lights {nl} svo {svo}")
                except AssertionError as e: bad+=1; print("MISMATCH",n,variant,cam,nl,svo,str(e)[:120],flush=True)
    print(n,variant,"done",flush=True)
print("bad",bad)
