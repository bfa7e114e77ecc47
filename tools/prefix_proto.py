"""Python model of the exact parallel float32 prefix chain used by the k-means++ pick kernel
(vecgo_b200/csrc/vg_kmeans.cu: pp_pick_parallel_kernel): per-chunk (K, c, t, parity) summaries, their composition, binade
crossings walked sequentially.  `par_prefix` must agree bit for bit with the sequential loop `seq` (the float32
running sum of pq.go:299-303) — tests/test_prefix_model_cpu.py fuzzes it; running this file does a longer fuzz."""
import numpy as np, struct
F=np.float32
def seq(x):
    s=F(0); out=np.empty(len(x),F)
    for i,v in enumerate(x):
        s=F(s+v); out[i]=s
    return out
def decomp(v):
    """float32 >= 0 -> (M, E) with v = M * 2^E, M < 2^24 integer"""
    b=struct.unpack('<I',struct.pack('<f',float(v)))[0]
    ex=(b>>23)&0xFF; fr=b&0x7FFFFF
    if ex==0: return fr, -149
    return fr|0x800000, ex-150
CH=32
HUGE=1<<40
def summarize(xs, Eu):
    """block summary (K, c, t, pi, huge) for elements xs under ulp 2^Eu, as a function of incoming integer s:
       F(s) = s + K + t*[(s + c) odd];  pi = parity of the output when t (tie seen)"""
    K=0; c=0; t=0; pi=0
    for v in xs:
        M,E=decomp(v)
        if M==0: continue
        d=Eu-E
        if d<=0:
            return (HUGE,0,0,0)   # certainly crosses
        if d>=25: continue
        fl=M>>d; rem=M&((1<<d)-1); half=1<<(d-1)
        if rem>half: m=fl+1; tie=False
        elif rem<half: m=fl; tie=False
        else: m=fl; tie=True
        if not tie:
            K+=m
            if t: pi=(pi+m)&1
        else:
            if not t:
                # first tie: r = s + K + fl ; result r + (r odd)  -> even
                # F(s) = s + K + fl + [(s + K + fl) odd]
                c=(K+fl)&1; K=K+fl; t=1; pi=0
            else:
                r_par=(pi+fl)&1
                K+=fl+r_par; pi=0
        if K>=HUGE: return (HUGE,0,0,0)
    return (K,c,t,pi)
def apply(summ, s):
    K,c,t,pi=summ
    return s+K+(t*((s+c)&1))
def compose(a,b):
    """b after a"""
    Ka,ca,ta,pa=a; Kb,cb,tb,pb=b
    if Ka>=HUGE or Kb>=HUGE: return (HUGE,0,0,0)
    if ta:
        # output parity of a known = pa
        extra=tb*((pa+cb)&1)
        K=Ka+Kb+extra
        if tb: pi=pb
        else: pi=(pa+Kb)&1
        return (K,ca,1,pi)
    else:
        if tb: return (Ka+Kb,(Ka+cb)&1,1,pb)
        return (Ka+Kb,0,0,0)
def par_prefix(x, nthreads=64):
    """returns total via the block algorithm (simulating rounds); also returns per-chunk start states for checking"""
    n=len(x); S=F(0); p=0
    starts={}
    while p<n:
        M,E=decomp(S)
        b=struct.unpack('<I',struct.pack('<f',float(S)))[0]
        ex=(b>>23)&0xFF
        if ex==0 or ex==255:
            # zero/denormal state: walk one chunk sequentially
            q=min(n,p+CH)
            starts[p]=S
            for v in x[p:q]: S=F(S+v)
            p=q; continue
        Eu=ex-150; s0=M  # S = s0 * 2^Eu, s0 in [2^23,2^24)
        W=min(n-p, nthreads*CH)
        nth=(W+CH-1)//CH
        summ=[summarize(x[p+i*CH:min(p+W,p+(i+1)*CH)],Eu) for i in range(nth)]
        # exclusive scan of compositions
        s_in=[]; acc=None
        cross=None
        s=s0
        for i in range(nth):
            # incoming state for thread i: apply composed prefix to s0 (sequential here; parallel scan in CUDA)
            s_i = s0 if acc is None else (None if acc[0]>=HUGE else apply(acc,s0))
            if s_i is None or s_i>=(1<<24):
                cross=i-1; break   # crossing happened inside thread i-1's chunk
            s_in.append(s_i)
            # does thread i's chunk cross?
            if summ[i][0]>=HUGE or s_i+summ[i][0]+1>=(1<<24):
                cross=i; break
            acc=summ[i] if acc is None else compose(acc,summ[i])
        if cross is None:
            for i in range(nth): starts[p+i*CH]=F(np.ldexp(float(s_in[i]),Eu))
            s_end=apply(acc,s0)
            S=F(np.ldexp(float(s_end),Eu)); p+=W
        else:
            for i in range(cross+1): starts[p+i*CH]=F(np.ldexp(float(s_in[i]),Eu))
            S=F(np.ldexp(float(s_in[cross]),Eu))
            a=p+cross*CH; q=min(p+W,a+CH)
            for v in x[a:q]: S=F(S+v)
            p=q
    return S, starts


def fuzz(trials, seed0=100):
    bad = 0
    for trial in range(trials):
        r = np.random.default_rng(trial + seed0)
        n = int(r.integers(1, 1500))
        kind = trial % 5
        if kind == 0:
            x = (r.random(n) * r.choice([1e-3, 1, 1e3])).astype(F)
        elif kind == 1:
            x = (2.0 ** r.integers(-12, 6, n)).astype(F)
        elif kind == 2:
            x = (r.integers(0, 9, n) * F(0.125)).astype(F)
        elif kind == 3:
            x = r.random(n).astype(F)
            idx = r.integers(0, n, max(1, n // 5))
            x[idx] = (2.0 ** r.integers(-24, 3, len(idx))).astype(F)
        else:
            x = (np.abs(r.standard_normal(n)) * 2.0 ** r.integers(-20, 20)).astype(F)
        ref = seq(x)
        tot, starts = par_prefix(x, nthreads=int(r.integers(1, 9)))
        ok = tot == ref[-1]
        for pos, val in starts.items():
            exp = F(0) if pos == 0 else ref[pos - 1]
            if val != exp:
                ok = False
        if not ok:
            bad += 1
    return bad


def special_cases():
    rng = np.random.default_rng(1)
    cases = {
        "uniform": (rng.random(5000) ** 2 * 10).astype(F),
        "pow2 (ties)": (2.0 ** rng.integers(-30, 10, 6000)).astype(F),
        "mixed range": np.concatenate([np.zeros(100, F), rng.random(3000).astype(F), np.full(500, F(1e-30)), (rng.random(2000) * 1e6).astype(F)]),
        "const 0.5 (ties)": np.full(9000, F(0.5)),
        "denormals": np.concatenate([np.full(10, F(1e-45)), np.full(3000, F(3e-39)), rng.random(3000).astype(F)]),
    }
    bad = []
    for name, x in cases.items():
        ref = seq(x)
        tot, starts = par_prefix(x)
        ok = tot == ref[-1] and all(val == (F(0) if pos == 0 else ref[pos - 1]) for pos, val in starts.items())
        if not ok:
            bad.append(name)
    return bad


if __name__ == "__main__":
    print("special cases failing:", special_cases())
    print("fuzz failures:", fuzz(400))
