import sys, os, importlib, time, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import _inputs as I
from oracle import pn2_oracle as O
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
DEV='cuda'
x_h = I.facade_batch(2, 2048, 9, 3).transpose(2, 1); x = x_h.to(DEV)
def rms(a, b): return ((a-b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt()).item()
for init in ('randomized', 'default'):
    torch.manual_seed(0)
    ref = O.OracleSemSeg(18, 3)
    if init == 'randomized': I.randomize_module_(ref, 61)
    ref.drop1.p = 0.0
    for mode in ('eval', 'train'):
        getattr(ref, mode)()
        torch.manual_seed(72); rp, _ = ref(x_h); rp = rp.detach()
        refg = __import__('copy').deepcopy(ref).to(DEV)
        torch.manual_seed(72); gp, _ = refg(x); gp = gp.detach().cpu()
        line = '%-10s %-5s | ref-GPU-fp32(TF32 conv): max %.2e rms %.2e ' % (
            init, mode, (gp-rp).abs().max(), rms(gp, rp))
        for prec in ('fp32', 'bf16'):
            pn2.set_precision(prec)
            net = pn2.get_model(18, 3); net.load_state_dict(ref.state_dict()); net.drop1.p = 0.0; net = net.to(DEV)
            getattr(net, mode)()
            torch.manual_seed(72); p, _ = net(x); p = p.detach().cpu()
            line += ' | ours-%s: max %.2e rms %.2e argmax %.4f' % (prec, (p-rp).abs().max(), rms(p, rp), (p.argmax(-1)==rp.argmax(-1)).float().mean())
        print(line)
# first timing look: config 2 train step
for prec in ('fp32', 'bf16'):
    pn2.set_precision(prec)
    net = pn2.get_model(18, 3).to(DEV).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    xb = I.facade_batch(32, 4096, 9, 11).to(DEV).transpose(2, 1)
    tgt = I.labels(32, 4096, 18, 3).to(DEV); w = torch.ones(18, device=DEV)
    def step():
        opt.zero_grad()
        pred, _ = net(xb)
        loss = torch.nn.functional.nll_loss(pred.contiguous().view(-1, 18), tgt, weight=w)
        loss.backward(); opt.step()
        return loss
    for _ in range(3): step()
    torch.cuda.synchronize(); l0 = pn2.launch_count(); t = time.time()
    for _ in range(10): step()
    torch.cuda.synchronize(); dt = (time.time()-t)/10
    print('%s train step 32x4096: %.2f ms  -> %.2f M points/s, %d pn2 launches/step' % (prec, dt*1e3, 32*4096/dt/1e6, (pn2.launch_count()-l0)//10))
    net.eval()
    with torch.no_grad():
        for _ in range(3): net(xb)
        torch.cuda.synchronize(); t = time.time()
        for _ in range(10): net(xb)
        torch.cuda.synchronize(); dt = (time.time()-t)/10
    print('%s eval fwd 32x4096: %.2f ms -> %.2f M points/s' % (prec, dt*1e3, 32*4096/dt/1e6))
