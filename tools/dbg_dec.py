import importlib, sys, torch
sys.path.insert(0, '/root/repo')
L = importlib.import_module("multi-modal-gnn_b200._lib"); lib = L.load()
dev = torch.device('cuda')
for m, frac in [(600,1.0),(128,1.0),(256,1.0),(1000,1.0),(43038,1.0)]:
    gen = torch.Generator().manual_seed(m + 1)
    n_p, n_l = 5000, 160
    U, V = torch.randn(n_p, 64, generator=gen).to(dev), torch.randn(n_l, 64, generator=gen).to(dev)
    pi, li = torch.randint(0, n_p, (m,), generator=gen).to(dev), torch.randint(0, n_l, (m,), generator=gen).to(dev)
    W2, b2 = (torch.randn(32, 64, generator=gen) / 8).to(dev), torch.randn(32, generator=gen).to(dev)
    w3 = torch.randn(32, generator=gen).to(dev)
    dpred = (torch.randn(m, generator=gen) * (torch.rand(m, generator=gen) < frac)).to(dev)
    ws = torch.empty(lib.b2g_decoder_bwd_ws_bytes(m), dtype=torch.uint8, device=dev)
    def run(fn):
        g = torch.zeros(m, 64, device=dev); flags = torch.empty(m, device=dev)
        dW2, db2, dw3, db3 = torch.empty(32, 64, device=dev), torch.empty(32, device=dev), torch.empty(32, device=dev), torch.empty(1, device=dev)
        L.check(fn(U.data_ptr(), V.data_ptr(), pi.data_ptr(), li.data_ptr(), W2.data_ptr(), b2.data_ptr(), w3.data_ptr(), dpred.data_ptr(), m, 0.0, 99, 1, 2, g.data_ptr(), flags.data_ptr(), dW2.data_ptr(), db2.data_ptr(), dw3.data_ptr(), db3.data_ptr(), ws.data_ptr(), ws.numel(), None))
        torch.cuda.synchronize(); return g*flags.unsqueeze(1), dW2, db2, dw3, db3
    r, o = run(lib.b2g_decoder_bwd), run(lib.b2g_decoder_bwd_tc)
    d = (r[0]-o[0]).abs(); rowerr = d.max(1)[0]; bad = (rowerr > 5e-3*r[0].abs().max()).nonzero().squeeze(1)
    print(m, 'g max err', d.max().item(), 'ref max', r[0].abs().max().item(), 'bad rows', bad.numel(), bad[:12].tolist(), 'dW2 err', (r[1]-o[1]).abs().max().item()/r[1].abs().max().item(), 'db2', (r[2]-o[2]).abs().max().item(), 'dw3', (r[3]-o[3]).abs().max().item())
    if bad.numel():
        i = int(bad[0]); print(' row', i, 'dy', dpred[i].item(), 'ref nz', (r[0][i]!=0).sum().item(), 'out nz', (o[0][i]!=0).sum().item(), r[0][i][:6].tolist(), o[0][i][:6].tolist())
