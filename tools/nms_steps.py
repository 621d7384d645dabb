import sys, torch, time
sys.path.insert(0, '/root/repo')
from glenet_b200 import iou3d_nms_utils as I, synth, _lib
dev = torch.device('cuda:0')
b, s = synth.proposals(4096, 20, 21)
b, s = b.to(dev), s.to(dev)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print('nms_gpu total          %.1f us' % t(lambda: I.nms_gpu(b, s, 0.7)))
print('sort                   %.1f us' % t(lambda: s.sort(0, descending=True)[1]))
order = s.sort(0, descending=True)[1]
print('gather boxes[order]    %.1f us' % t(lambda: b[order].contiguous()))
bs = b[order].contiguous().unsqueeze(0)
print('_nms_sorted (2 kernels + allocs) %.1f us' % t(lambda: I._nms_sorted("glenet_nms_gpu", bs, 0.7)))
keep, num = I._nms_sorted("glenet_nms_gpu", bs, 0.7)
print('num.item()             %.1f us' % t(lambda: int(num.item())))
n = int(num.item())
print('order[keep[:n]]        %.1f us' % t(lambda: order[keep[0, :n]].contiguous()))
lib = _lib.load()
ws_bytes = lib.glenet_nms_workspace_bytes(1, 4096)
ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
k2 = torch.empty((1, 4096), dtype=torch.int64, device=dev); n2 = torch.zeros((1,), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
print('C call only (2 kernels) %.1f us' % t(lambda: lib.glenet_nms_gpu(bs.data_ptr(), 1, 4096, 0.7, k2.data_ptr(), n2.data_ptr(), ws.data_ptr(), ws_bytes, st)))
print('kept', n)
