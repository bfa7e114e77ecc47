import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import vecgo_b200 as vg
L = vg._lib
F = np.float32
def run(n, dim, nq, k, maskfrac=None, seed=11):
    rng = np.random.default_rng(seed)
    x = rng.random((n, dim)).astype(F); q = rng.random((nq, dim)).astype(F)
    dev = torch.device("cuda:0")
    L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
    ix = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n); ix.upload(vectors=x)
    dq = torch.from_numpy(q).to(dev)
    dm = 0
    if maskfrac is not None:
        keep = rng.random(n) < maskfrac
        mask = np.packbits(keep, bitorder="little")
        tm = torch.from_numpy(np.concatenate([mask, np.zeros(8, np.uint8)])).to(dev); dm = tm.data_ptr()
    r = torch.empty((nq, k), dtype=torch.int32, device=dev); s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    c = torch.empty((nq,), dtype=torch.int32, device=dev); f = torch.zeros((nq,), dtype=torch.int32, device=dev)
    l0 = vg.launch_count()
    ix.search_dev_async(dq.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr(), f.data_ptr(), d_mask=dm)
    torch.cuda.synchronize()
    nl = vg.launch_count() - l0
    r1, s1, f1 = r.cpu().numpy().view(np.uint32), s.cpu().numpy(), f.cpu().numpy()
    L.call("vg_flat_tc_enable", 0)
    ix.search_dev(dq.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr(), d_mask=dm)
    L.call("vg_flat_tc_enable", 1)
    r2, s2 = r.cpu().numpy().view(np.uint32), s.cpu().numpy()
    bad = [i for i in range(nq) if not f1[i] and not (np.array_equal(r1[i], r2[i]) and np.array_equal(s1[i].view(np.uint32), s2[i].view(np.uint32)))]
    print(f"n={n} dim={dim} nq={nq} k={k} mask={maskfrac}: launches {nl}, flags {int((f1 != 0).sum())} codes {np.bincount(f1, minlength=6).tolist()}, WRONG unflagged {len(bad)}")
    for i in bad[:2]:
        print("  q", i, "got", r1[i], s1[i]); print("      want", r2[i], s2[i])
    ix.close()
for a in [(100000,128,1000,10,None),(9000,96,50,10,0.4),(20000,128,256,10,0.5),(8192,128,256,10,None),(40000,32,600,10,None),(300000,256,512,16,None),(50000,64,100,5,0.05)]:
    run(*a)
