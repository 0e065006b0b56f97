"""Run the forward pass of X3D-M (train-mode BN) twice on the same clip batch and report whether the logits agree bit for bit."""
import sys, torch
sys.path.insert(0, '/root/repo')
import x3d_multigrid_b200 as X
torch.manual_seed(0)
for dtype in (torch.bfloat16, torch.float32):
    m = X.generate_model('M', n_classes=400, base_bn_splits=2, dropout=0.0).cuda().set_compute_dtype(dtype).train()
    x = torch.randn(8, 3, 16, 160, 160, device='cuda')
    with torch.no_grad():
        a = m(x).clone(); b = m(x).clone(); c = m(x).clone()
    torch.cuda.synchronize()
    print(dtype, 'bitwise equal:', bool(torch.equal(a, b) and torch.equal(a, c)),
          'max |diff| / max |logit|:', float((a - b).abs().max() / a.abs().max()))
