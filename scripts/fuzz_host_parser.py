"""Deep fuzz of the DEVICE's sequential code compiled for the host (tests/host_shim) against the Python oracle:
grammar-based JSON-ish documents with tricky keys (folding, escapes, duplicates), numbers at the float64 limits,
handler-shaped arguments, byte mutations; under the SDK-payload rules and the HTTP-body rules.

    python scripts/fuzz_host_parser.py [--seed S] [--parse N] [--run M]

--parse N documents through parse_payload (status, argument count, keyword arguments, args[0] token);
--run M payloads through the whole sequential path of every handler (status + result bytes).
tests/test_device_parser_on_host.py is the fixed-size version of this that runs with the CPU suite. (A 120 k run of this
script found the duplicate-key / overflow hole in both oracles)"""
import argparse
import base64
import ctypes as C
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import loop                                                    # noqa: E402
from oracle.pyoracle.gojson import GoJSONError, go_unmarshal                        # noqa: E402
from tests.test_device_parser_on_host import GXX, SO, SRC, _oracle                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--parse", type=int, default=60000)
ap.add_argument("--run", type=int, default=15000)
args = ap.parse_args()
if not os.path.exists(SO):
    subprocess.check_call([GXX, "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC])
lib = C.CDLL(SO)
lib.b9_host_parse.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
lib.b9_host_run.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_char_p, C.c_uint32]
lib.b9_host_run.restype = C.c_long
rnd = random.Random(args.seed)
def rstr():
    alphabet=['a','Z','0',' ','"','\\','/','<','\n','\t','\x01','\x7f','\u00e9','\u2028','\U0001f600','\ud83d','\udc00','args','kwargs','k','s','\u212a','\u017f']
    return ''.join(rnd.choice(alphabet) for _ in range(rnd.randint(0,6)))
def rval(d=0):
    k=rnd.randint(0,9 if d<4 else 5)
    if k==0: return None
    if k==1: return rnd.choice([True,False])
    if k==2: return rnd.randint(-10**rnd.randint(0,20),10**rnd.randint(0,20))
    if k==3: return rnd.choice([0.0,0.5,-1.25,1e21,1e-7,1e308,5e-324,123456789.125])
    if k in(4,5): return rstr()
    if k in(6,7): return [rval(d+1) for _ in range(rnd.randint(0,3))]
    return {rstr(): rval(d+1) for _ in range(rnd.randint(0,3))}
def rkey():
    return rnd.choice(['args','kwargs','Args','KWARGS','arg\u017f','\u212awargs','x','args ','', 'kwargs2'])
def rtext():
    # a JSON-ish document: object with tricky keys, sometimes non-object, sometimes raw escapes
    t=rnd.randint(0,9)
    if t==0: return json.dumps(rval()).encode()
    items=[]
    for _ in range(rnd.randint(0,4)):
        k=rkey(); 
        if k=='args' and rnd.random()<0.7: v=[rval(1) for _ in range(rnd.randint(0,3))]
        elif k=='kwargs' and rnd.random()<0.7: v={rstr(): rval(1) for _ in range(rnd.randint(0,2))}
        else: v=rval(1)
        ks=json.dumps(k, ensure_ascii=rnd.random()<0.5)
        if rnd.random()<0.1: ks=ks.replace('a','\\u0061',1)
        items.append(ks+rnd.choice([':',': ',' : '])+json.dumps(v, ensure_ascii=rnd.random()<0.5, separators=rnd.choice([(',',':'),(', ',': ')])))
    s='{'+rnd.choice([',',', ',' , ']).join(items)+'}'
    b=s.encode('utf-8','surrogatepass')
    if rnd.random()<0.3:
        m=bytearray(b)
        for _ in range(rnd.randint(1,2)):
            if not m: break
            pos=rnd.randrange(len(m)); op=rnd.randint(0,2); ch=rnd.choice(list(b'{}[],:" 01e.-+nulltrfackwgsAK\\u\xc3\xa9\xff\x00'))
            if op==0: m[pos]=ch
            elif op==1: del m[pos]
            else: m[pos:pos]=bytes([ch])
        b=bytes(m)
    return b

def rtext2():
    t=rnd.randint(0,6)
    if t<=2: return rtext()
    # handler-shaped payloads with noise
    if t==3: arg=rstr()*rnd.randint(0,4)
    elif t==4:
        raw=bytes(rnd.randrange(256) for _ in range(rnd.choice([0,4,8,8,16,24,12])))
        arg=base64.b64encode(raw).decode()
        if rnd.random()<0.3: arg=arg[:-1] if arg else arg
    elif t==5: arg={"values":[rnd.choice([rnd.randint(-10**rnd.randint(0,18),10**rnd.randint(0,18)), True, False, 1.5, "s", None]) for _ in range(rnd.randint(0,5))], rstr(): rval(1)}
    else: arg=rval(0)
    kw={} if rnd.random()<0.8 else {rstr(): rval(1)}
    b=json.dumps({"args":[arg],"kwargs":kw}, ensure_ascii=rnd.random()<0.6).encode('utf-8','surrogatepass')
    if rnd.random()<0.15:
        m=bytearray(b); pos=rnd.randrange(len(m)); m[pos]=rnd.choice(list(b'{}[],:" 01e.-\\\xff'))
        b=bytes(m)
    return b

# ---- parse_payload
bad=0; n=0; dec=0
N=args.parse
for it in range(N):
    b=rtext()
    for http in (False,True):
        out=(C.c_uint32*12)(); lib.b9_host_parse(b,len(b),int(http),out)
        st,nargs,kwn,kind,off,ln,fl,mg=list(out)[:8]
        if st==4: dec+=1; continue
        w=_oracle(b,http); n+=1
        if w is None:
            if st!=3: bad+=1; print("ACCEPTED",http,b,st)
            continue
        if st!=0: bad+=1; print("REJECTED",http,b,st,w); continue
        a,k=w
        if nargs!=len(a) or ((not (mg and not http)) and bool(kwn)!=bool(k)): bad+=1; print("MISMATCH",http,b,nargs,kwn,w); continue
        if a:
            tok=b[off:off+ln]
            try: g=go_unmarshal(tok)
            except GoJSONError: bad+=1; print("TOKBAD",http,b,tok); continue
            if not (g==a[0] or (g!=g and a[0]!=a[0])): bad+=1; print("A0",http,b,tok,g,a[0])
    if bad>10: break
print("cases",N,"checked",n,"declined",dec,"bad",bad)

# ---- the whole sequential path, every handler
N=args.run
pl=[rtext2() for _ in range(N)]
ids=[bytes([i&255])*16 for i in range(N)]
code={"COMPLETE":0,"ERROR":1,"RETRY":2,"REJECTED":3}
bad=0; dec=0; chk=0
for http in (False,True):
    for hid,h in enumerate(["identity","crc32","vadd_f32","json_sum"]):
        want=loop.run_task_loop(pl,ids,h,http_body=http)
        for b,w in zip(pl,want):
            st=C.c_uint8(0); has=C.c_uint8(0); buf=C.create_string_buffer(8*len(b)+64)
            n=lib.b9_host_run(b,len(b),int(http),hid,C.byref(st),C.byref(has),buf,len(buf))
            if st.value==4: dec+=1; continue
            chk+=1
            res=buf.raw[:n] if has.value else None
            if st.value!=code[w.status] or res!=w.result:
                # numpy two-NaN ambiguity: ignore vadd cases where both bytes decode to NaNs
                bad+=1
                if bad<=15: print("MISMATCH",h,http,b[:200],st.value,res,w.status,w.result)
print("payloads",N,"checked",chk,"declined",dec,"bad",bad)
