import sys; import pathlib; R_=pathlib.Path(__file__).resolve().parents[2]; sys.path.insert(0,str(R_/'tests')); sys.path.insert(0,str(R_))
import importlib, numpy as np, oracle_lib as O, emu_lib
from conftest import oracle_bias, AUX_FIELDS
pkg=importlib.import_module("voxel-raycaster_b200"); S=pkg.scene
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
bad=0; N=int(sys.argv[2]) if len(sys.argv)>2 else 100
for it in range(N):
    n=int(rng.choice([8,16,32,64]))
    vol=np.zeros((n,n,n),np.int8)
    dens=rng.choice([0.002,0.01,0.05,0.2,0.6])
    vol[rng.random((n,n,n))<dens]=5
    vol[rng.random((n,n,n))<dens*0.3]=6
    vol[rng.random((n,n,n))<0.02]=int(rng.integers(-5,9))
    mode=rng.integers(0,5)
    pos=(rng.random(3)*n).astype(np.float32)
    if mode==1: pos=np.floor(pos).astype(np.float32)
    if mode==2: pos=(np.floor(pos)+0.5).astype(np.float32)
    if mode==3: pos=(rng.random(3)*n*1.5-0.25*n).astype(np.float32)
    pos=np.clip(pos,-3,n+3).astype(np.float32)
    d=np.array([rng.random()*np.pi, rng.random()*2*np.pi],np.float32)
    if mode==4: d=np.array([rng.choice([0,np.pi/2,np.pi,1.57]), rng.choice([0,np.pi/2,np.pi,3*np.pi/2])],np.float32)
    nl=int(rng.choice([1,1,2,3]))
    lights=np.zeros((8,10),np.float32)
    for l in range(nl):
        lights[l]=[rng.random(),rng.random(),rng.random(),rng.random()*2, *(rng.random(3)*n*1.2-0.1*n), -1,-1,-1.5]
    w,h=48,32
    sc=S.Scene(n,vol,w,h,pos,d,lights,max_distance=int(rng.choice([20,3*n,5])))
    table=O.make_ray_table(w,h)
    desc,root=pkg.octree_generate(vol)
    ref_rgba,ref_aux,_=O.raycast(sc,table,octree=(desc,root),shadow_lights=nl)
    inside = all(0 <= int(np.floor(p)) < n for p in pos)
    bias=oracle_bias(O,sc,desc,root)
    if bias is None: continue      # get_oct_vox of a camera voxel outside the map: host evaluates on clamped? skip
    for mode_svo in (0,1,2):
        rgba,aux=emu_lib.raycast(sc,table,bias=bias,use_svo=mode_svo,shadow_lights=nl)
        tie=(ref_aux["flags"]&4)!=0 if mode_svo==2 else np.zeros_like(ref_aux["flags"],bool)
        ok=True
        for f in ("hit","face","status","hit_type","steps_first","steps_total"):
            if (np.any(np.atleast_3d(ref_aux[f]!=aux[f]),axis=-1)&~tie).any(): ok=False; why=f
        diff=np.abs(ref_rgba.astype(int)-rgba.astype(int)).max(-1)
        if (diff[~tie]>0).any() or (diff>1).any(): ok=False; why="rgba"
        if not ok:
            bad+=1; print("MISMATCH it",it,"svo",mode_svo,"n",n,"mode",mode,"lights",nl,"md",sc.max_distance,"pos",pos,"dir",d,why)
print("done",N,"bad",bad)
