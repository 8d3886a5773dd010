"""Counts the distinct 128 B lines (and 32 B sectors) one 32-lane gather of the tet pass touches, for the x-fastest volume and
for bricked layouts, and for 8 / 16 / 32 lanes per tet, on a slab of the C4 benchmark mesh (numpy only).  The figure for the
linear layout with 8 lanes per tet (16.8) matches ncu (L1 tag requests per gather of tet_integrate_w2_kernel)."""
import sys, importlib.util, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, ROOT)
import gen_tables
spec = importlib.util.spec_from_file_location("syn", os.path.join(ROOT, "3d-image-segmentation-with-mesh_b200", "synthetic.py")); syn = importlib.util.module_from_spec(spec); spec.loader.exec_module(syn)
tet_tab,_ = gen_tables.build()
tabs=[np.array(t) for t in tet_tab]
n=512; edge=syn.edge_for_cells(n,69)
pos,tets,grid=syn.grid_mesh((n,n,n),edge,0.2,1234)
base=syn.boundary_layer_labels(tets,grid)
# take a slab of consecutive tets in the middle
start=(len(tets)//2//32)*32
T=tets[start:start+32*600]; B=base[start:start+32*600]
P=pos[T]  # [m,4,3]
a,b,c,d=P[:,0],P[:,1],P[:,2],P[:,3]
vol=np.abs(np.einsum('ij,ij->i',a-d,np.cross(b-d,c-d))/6)
lvl=np.clip(np.floor(np.cbrt(vol)+1e-9).astype(int),1,9)
print('levels',np.bincount(lvl), 'labelled',(B!=0).mean())
def samples(t):
    tb=tabs[lvl[t]-1]
    return np.floor(tb@P[t]).astype(np.int64)  # [n3,3]
def line_ids(v,layout):
    x,y,z=v[:,0],v[:,1],v[:,2]
    if layout=='lin': return (z*n+y)*(n//32)+x//32, (z*n+y)*(n//8)+x//8
    bx,by,bz=layout
    # brick of bx*by*bz voxels = 32 voxels (128B); sector = 8 voxels: assume x-fastest inside brick
    lid=((z//bz)*(n//by)+(y//by))*(n//bx)+(x//bx)
    inner=((z%bz)*by+(y%by))*bx+(x%bx)
    return lid, lid*4+inner//8
def sim(mapping,layout,order=None):
    lines=0;sect=0;instr=0;lanes=0
    for tile in range(len(T)//32):
        S=[samples(tile*32+k) for k in range(32)]
        if order is not None: S=[s[order[len(s)]] for s in S]
        if mapping=='g8':
            for r in range(8):
                ts=[S[4*r+g] for g in range(4)]
                nmax=max(len(s) for s in ts)
                for basei in range(0,nmax,32):
                    for u in range(4):
                        vs=[s[basei+8*u:basei+8*u+8] for s in ts if len(s)>basei+8*u]
                        if not vs: continue
                        v=np.concatenate(vs)
                        l,s_=line_ids(v,layout)
                        lines+=len(np.unique(l)); sect+=len(np.unique(s_)); instr+=1; lanes+=len(v)
        elif mapping=='g32':
            for k in range(32):
                s=S[k]
                for basei in range(0,len(s),32):
                    v=s[basei:basei+32]
                    l,s_=line_ids(v,layout)
                    lines+=len(np.unique(l)); sect+=len(np.unique(s_)); instr+=1; lanes+=len(v)
        elif mapping=='g16':
            for r in range(16):
                ts=[S[2*r+g] for g in range(2)]
                nmax=max(len(s) for s in ts)
                for basei in range(0,nmax,64):
                    for u in range(4):
                        vs=[s[basei+16*u:basei+16*u+16] for s in ts if len(s)>basei+16*u]
                        if not vs: continue
                        v=np.concatenate(vs)
                        l,s_=line_ids(v,layout)
                        lines+=len(np.unique(l)); sect+=len(np.unique(s_)); instr+=1; lanes+=len(v)
    return dict(mapping=mapping,layout=layout,lines_per_instr=round(lines/instr,2),sect_per_instr=round(sect/instr,2),lanes=round(lanes/instr,1),lines_per_sample=round(lines/lanes,3), total_lines=lines)
# spatial ordering of table rows: sort rows by a Morton-ish key of barycentric position
def morton_order(tb):
    q=np.floor(tb[:,1:4]*16).astype(int)
    key=np.zeros(len(tb),int)
    for bit in range(4):
        for k in range(3):
            key|=((q[:,k]>>bit)&1)<<(3*bit+k)
    return np.argsort(key,kind='stable')
order={len(t):morton_order(t) for t in tabs}
for layout in ['lin',(4,4,2),(8,2,2),(2,4,4),(4,2,4)]:
    for mapping in ['g8','g16','g32']:
        print(sim(mapping,layout))
    print('morton rows', sim('g8',layout,order), sim('g32',layout,order))
