"""32-lane emulation of `warp_string_fast` (simdjson-go_b200/csrc/stage2.cuh) against the CPU oracle.

The CUDA routine decodes every escape of a 32-byte window at once: run-parity escape starts, neighbour
shuffles for the hex digits, consumed-range masks, prefix counts for the output positions, decoding limited
to starts at lanes <= 20, exact single step for chains of high surrogates.  This file restates exactly that
algorithm with 32-element Python lists standing in for the lanes (ballot = bit mask over the list,
shfl_down = index + k) and checks validity, source length and unescaped bytes against the oracle's
parse_string (parse_string_amd64.s:72-479) on random strings built from escape fragments, quotes inside
escapes, surrogate halves and chains.  It is how the algorithm was verified before any GPU time was spent
(80 000 strings then; a few thousand here to keep the CPU suite short); on the GPU the test hook
sj_test_parse_strings runs the real kernel code against the exact routines on every input.  CPU only."""
import numpy as np

M32=0xffffffff
def at(body,i): return body[i] if i < len(body) else 0
def d2v(c):
    if c < 0x30: return 0
    if c <= 0x39: return c-0x30
    l=c|0x20
    if c<0x80 and 0x61<=l<=0x66 and c>=0x41: return l-0x61+10
    return -1
EM={0x22:0x22,0x2f:0x2f,0x5c:0x5c,ord('b'):8,ord('f'):12,ord('n'):10,ord('r'):13,ord('t'):9}
def ballot(pred): 
    m=0
    for l in range(32):
        if pred[l]: m|=1<<l
    return m
def ffs(x): return (x & -x).bit_length()  # 1-based, 0 if none
def clz(x): return 32 - x.bit_length()
def enc(cp,n):
    if n==1: return bytes([cp&0xff])
    if n==2: return bytes([(0xC0+(cp>>6))&0xff,0x80|(cp&63)])
    if n==3: return bytes([(0xE0+(cp>>12))&0xff,0x80|((cp>>6)&63),0x80|(cp&63)])
    return bytes([(0xF0+(cp>>18))&0xff,0x80|((cp>>12)&63),0x80|((cp>>6)&63),0x80|(cp&63)])

def window(body,p):
    """returns dict: status in {'bad','end','cont','fallback'}; consumed src E; out bytes"""
    c=[at(body,p+l) for l in range(32)]
    bs=ballot([x==0x5c for x in c]); qm=ballot([x==0x22 for x in c]); um=ballot([x==ord('u') for x in c])
    if bs==0:
        if qm==0: return dict(status='cont',E=32,out=bytes(c))
        Q=ffs(qm)-1
        return dict(status='end',E=Q,out=bytes(c[:Q]))
    # parity starts
    sp=[False]*32
    for L in range(32):
        if (bs>>L)&1:
            below=(~bs)&((1<<L)-1)&M32
            R=(32-clz(below)) if below else 0
            sp[L]=((L-R)&1)==0
    SP=ballot(sp)
    dv=[d2v(x) for x in c]
    def down(arr,L,k,default=0): return arr[L+k] if L+k<32 else default
    # per-lane 'u' decode (every lane computes as if it were a start; only parity-starts matter)
    e=[down(c,L,1) for L in range(32)]
    isu=[e[L]==ord('u') for L in range(32)]
    cp=[0]*32; uok=[False]*32
    for L in range(32):
        d2,d3,d4,d5=[down(dv,L,k,-1) for k in (2,3,4,5)]
        x=((d2<<12)&M32)|((d3<<8)&M32)|((d4<<4)&M32)|(d5&M32)
        cp[L]=x&M32
        noq=((qm>>(L+2))&0xF)==0
        uok[L]=noq and cp[L]<=0xFFFF and L+5<32
    high=[sp[L] and isu[L] and uok[L] and (cp[L]&0xFC00)==0xD800 for L in range(32)]
    H=ballot(high)
    if H & ((H<<6)&M32): return dict(status='fallback')
    lowc=(H<<6)&M32
    real=SP & ~lowc & M32
    adv=[0]*32; n=[0]*32; ok=[True]*32; outcp=[0]*32
    for L in range(32):
        if not ((real>>L)&1): continue
        if L>20: continue  # late start: not processed
        if not isu[L]:
            m=EM.get(e[L],0)
            if m==0: ok[L]=False
            adv[L]=2; n[L]=1; outcp[L]=m
        else:
            if not uok[L]: ok[L]=False; adv[L]=6; n[L]=1; continue
            x=cp[L]
            if (x&0xFC00)==0xD800:
                # pair: L+6 must be '\\' , L+7 'u', cp2 from lane L+6's own decode
                pair_ok=((bs>>(L+6))&1)==1 and ((um>>(L+7))&1)==1 and uok[L+6]
                if not pair_ok: ok[L]=False; adv[L]=12; n[L]=4; continue
                x2=cp[L+6]
                x=((((x<<10)+0xFCA00000)&M32 | ((x2+0xFFFF2400)&M32)) + 0x10000)&M32
                adv[L]=12
                if x>0x10FFFF: ok[L]=False; n[L]=4; continue
                n[L]=1 if x<0x80 else 2 if x<0x800 else 3 if x<0x10000 else 4
                outcp[L]=x
            else:
                adv[L]=6
                n[L]=1 if x<0x80 else 2 if x<0x800 else 3
                outcp[L]=x
    C=0
    for L in range(21):
        if (real>>L)&1: C|=(((1<<adv[L])-1)<<L)&M32
    late=SP & ~C & ~((1<<21)-1) & M32
    Z=ffs(late)-1 if late else 32
    qreal=qm & ~C & M32
    Q=ffs(qreal)-1 if qreal else 32
    E=min(Q,Z)
    # validity of every real processed start inside the region
    for L in range(21):
        if (real>>L)&1 and L<E and not ok[L]: return dict(status='bad')
    out=bytearray()
    for L in range(E):
        if (real>>L)&1 and L<=20: out+=enc(outcp[L],n[L])
        elif not ((C>>L)&1): out.append(c[L])
    return dict(status='end' if Q<Z else 'cont',E=E,out=bytes(out))

def fast(body):
    """returns None (invalid) or (src_len, out bytes); fallback = exact serial step"""
    p=0; out=bytearray()
    guard=0
    while True:
        guard+=1
        assert guard<10000
        if p>=len(body)+32: return None   # ran off the end: the exact routine decides (false)
        w=window(body,p)
        if w['status']=='fallback':
            # one exact serial step
            win=[at(body,p+l) for l in range(32)]
            ev=[l for l in range(32) if win[l] in (0x22,0x5c)]
            j=ev[0]
            out+=bytes(win[:j])
            if win[j]==0x22: return p+j,bytes(out)
            r=escape_step(body,p+j)
            if r is None: return None
            a,cpx,nn=r; out+=enc(cpx,nn); p+=j+a
            continue
        if w['status']=='bad': return None
        out+=w['out']
        if w['status']=='end': return p+w['E'],bytes(out)
        p+=w['E']
def escape_step(body,b):
    e=at(body,b+1)
    if e!=ord('u'):
        m=EM.get(e,0)
        if m==0: return None
        return 2,m,1
    def u(x): return x & M32
    c=[at(body,b+k) for k in range(12)]
    if 0x22 in c[2:6]: return None
    cp=(u(d2v(c[2])<<12)|u(d2v(c[3])<<8)|u(d2v(c[4])<<4)|u(d2v(c[5])))&M32
    a=6
    if (cp&0xFFFFFC00)==0xD800:
        if c[6]!=0x5c or c[7]!=ord('u'): return None
        if 0x22 in c[8:12]: return None
        cp2=(u(d2v(c[8])<<12)|u(d2v(c[9])<<8)|u(d2v(c[10])<<4)|u(d2v(c[11])))&M32
        if (cp|cp2)>0xFFFF: return None
        cp=((((cp<<10)+0xFCA00000)&M32 | ((cp2+0xFFFF2400)&M32)) + 0x10000)&M32
        a=12
    if cp<0x80:n=1
    elif cp<0x800:n=2
    elif cp<0x10000:n=3
    elif cp<=0x10FFFF:n=4
    else: return None
    return a,cp,n



ALPHABET = [b"a", b"\\", b'"', b"u", b"d", b"8", b"0", b"F", b"c", b"\\u", b"\\ud83d", b"\\ude00", b"n", b"/", b"\x00", b"-",
            b"\xc3\xa9", b"\\\\", b'\\"', b"xyz" * 5, b"t" * 31, b"q" * 33, b"\\u00e9", b"\\n", b" ", b"\\uD800", b"\\u1", b"\\u12",
            b"\\ud83d\\u", b"\\ud83d\\ud8", b"\\udbff\\u00", b"\\ud800\\ud800", b"!", b"\\ud83d\\n", b"\\u30c6", b"\\u30b9\\u30c8",
            b"\\\\\\\\\\\\", b"\\/"]


def test_lane_parallel_escape_windows_match_the_oracle(oracle):
    rng = np.random.default_rng(11)
    nvalid = 0
    for _ in range(20000):
        body = b"".join(ALPHABET[j] for j in rng.integers(0, len(ALPHABET), rng.integers(0, 30)))
        tail = b'"' if rng.integers(0, 10) else b""
        it = b'"' + body + tail + b"," * int(rng.integers(0, 3))
        ok_o, sl_o, dl_o = oracle.parse_string_validate_only(it, len(it) + 40)
        r = fast(it[1:])
        assert (r is not None) == ok_o, it
        if ok_o:
            nvalid += 1
            assert r[0] == sl_o and r[1] == oracle.parse_string(it)[1], it
    assert nvalid > 2000
