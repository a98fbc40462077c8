import importlib, sys, torch
sys.path.insert(0, '/root/repo')
L = importlib.import_module("multi-modal-gnn_b200._lib"); lib = L.load()
dev = torch.device('cuda')
for (m,n,k) in [(64,128,128),(64,256,128),(64,128,32),(64,32,128),(64,64,128),(128,128,128)]:
    g = torch.Generator().manual_seed(1)
    dy = (torch.randint(-8,9,(m,n),generator=g).float()/8).to(dev)
    x = (torch.randint(-8,9,(m,k),generator=g).float()/16).to(dev)
    ref = dy.double().t() @ x.double()
    dw = torch.full((n,k), float('nan'), device=dev)
    ws = torch.full((lib.b2g_linear_bwd_weight_tc_ws_bytes(m,n,k)//4,), 7.0, dtype=torch.float32, device=dev)
    rc = lib.b2g_linear_bwd_weight_tc(dy.data_ptr(), x.data_ptr(), m, n, k, dw.data_ptr(), ws.data_ptr(), ws.numel()*4, None)
    torch.cuda.synchronize()
    part = ws[:128*(n if n!=128 else k)]
    err = (dw.double()-ref).abs().max().item()
    print((m,n,k), 'rc', rc, 'err', err, 'partial nonzero frac', (part!=0).float().mean().item(), 'partial==7 frac', (part==7).float().mean().item(), 'dw nan', dw.isnan().any().item(), 'ref max', ref.abs().max().item(), 'dw max', dw.abs().max().item())
