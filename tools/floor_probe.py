import sys; sys.path.insert(0,'/root/repo')
from ecc_ldpc_b200 import codes, lib
lib.init(0)
code=codes.load_builtin('1920.1280.3.303'); dc=lib.DeviceCode(code)
for prec in ('fp32','fp64'):
    dec=lib.Decoder(dc,'spa',prec,max_batch=16384)
    for eb in (2.5,3.0):
        c=dec.simulate(50,eb,code.K,code.N,7,0,200000)
        print(prec,eb,c['frame_errors'],c['bit_errors'],c['unconverged'],c['sum_iters']/c['frames'],flush=True)
