import sys; import pathlib; R_=pathlib.Path(__file__).resolve().parents[2]; sys.path.insert(0,str(R_/'tests')); sys.path.insert(0,str(R_))
import importlib, numpy as np, oracle_lib as O, ref_kernel_lib as R
pkg=importlib.import_module("voxel-raycaster_b200"); S=pkg.scene
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
bad=0; N=int(sys.argv[2]) if len(sys.argv)>2 else 200
for it in range(N):
    n=int(rng.choice([8,16,32]))
    vol=np.zeros((n,n,n),np.int8)
    dens=rng.choice([0.01,0.05,0.2,0.6])
    vol[rng.random((n,n,n))<dens]=5
    vol[rng.random((n,n,n))<dens*0.3]=6
    vol[rng.random((n,n,n))<0.02]=int(rng.integers(-5,9))
    mode=rng.integers(0,5)
    pos=(rng.random(3)*n).astype(np.float32)
    if mode==1: pos=np.floor(pos).astype(np.float32)            # integer coordinates (frac 0)
    if mode==2: pos=(np.floor(pos)+0.5).astype(np.float32)
    if mode==3: pos=(rng.random(3)*n*1.5-0.25*n).astype(np.float32)   # possibly outside the map
    pos=np.clip(pos,-3,n+3).astype(np.float32)
    d=np.array([rng.random()*np.pi, rng.random()*2*np.pi],np.float32)
    if mode==4: d=np.array([rng.choice([0,np.pi/2,np.pi,1.57]), rng.choice([0,np.pi/2,np.pi,3*np.pi/2])],np.float32)
    lights=np.zeros((8,10),np.float32)
    lights[0]=[rng.random(),rng.random(),rng.random(),rng.random()*2, *(rng.random(3)*n*1.2-0.1*n), -1,-1,-1.5]
    if rng.random()<0.1: lights[0,4:7]=np.floor(lights[0,4:7])
    w,h=48,32
    sc=S.Scene(n,vol,w,h,pos,d,lights,max_distance=int(rng.choice([20,3*n,5])))
    # camera voxel must be inside the map for get_oct_vox / map indexing in the kernel? (the reference reads map only after the bounds test)
    desc,root=pkg.octree_generate(vol)
    try:
        o_rgba,o_aux,_=O.raycast(sc,octree=(desc,root))
        r_rgba,wr=R.raycast(sc,octree=(desc,root),lifted=True)
    except Exception as e:
        print("EXC",it,e); bad+=1; continue
    diff=np.abs(o_rgba.astype(int)-r_rgba.astype(int)).max(-1)
    skipped=(o_aux["status"]==0)|(o_aux["status"]==4)
    if diff.any() or not np.array_equal(wr,~skipped):
        bad+=1
        ys,xs=np.nonzero((diff>0)|(wr==skipped))
        print("MISMATCH it",it,"n",n,"mode",mode,"pos",pos,"dir",d,"md",sc.max_distance,"px",len(ys),"first",ys[0],xs[0],o_rgba[ys[0],xs[0]],r_rgba[ys[0],xs[0]],o_aux[ys[0],xs[0]], "written",wr[ys[0],xs[0]])
print("done",N,"bad",bad)
