"""Multi-GPU sanity (torchrun, NCCL): measurement-parallel tensor-network mode and slice-parallel amplitudes give
the single-rank numbers.  Prints OK lines on rank 0."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
import tedq_b200 as qb
from tedq_b200 import workloads as W

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank()
# 1. measurements dealt to ranks (C3-shaped: HEA with one Z expval per qubit)
spec = W.hea(8, 3)
circ = W.build_circuit(spec, qb)
x = torch.tensor(np.random.RandomState(0).rand(5, spec["n_params"]), dtype=torch.float32, device="cuda")
ref = circ.compilecircuit(backend="pytorch_b200").batched(x)
cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, hyper_opt={"max_repeats": 2, "measurement_parallel": True})
got = cc.batched(x)
err1 = float((got - ref).abs().max())
# 2. slices dealt to ranks, one all-reduce (C5-shaped, small lattice)
spec = W.lattice_rcs(4, 4, 6, seed=2, measure="state")
circ = W.build_circuit(spec, qb)
bits = [1, 0] * 8
one = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False,
                          hyper_opt={"max_repeats": 4, "slicing_opts": {"target_num_slices": 8}})
par = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False,
                          hyper_opt={"max_repeats": 4, "slicing_opts": {"target_num_slices": 8, "contract_parallel": True}})
a1, a2 = complex(one.amplitude(bits).cpu()), complex(par.amplitude(bits).cpu())
err2 = abs(a1 - a2) / abs(a1)
# 3. slice-sharded reverse pass: forward + reverse pass of this rank's slices, one all-reduce of the gradients
spec = W.mbl_1d(8)
circ = W.build_circuit(spec, qb)
x = torch.tensor(W.c2_inputs(3, 8, 1), device="cuda")
res = []
for par_on in (False, True):
    cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False,
                             hyper_opt={"max_repeats": 2, "tn_backward": "tree",
                                        "slicing_opts": {"target_num_slices": 8, "contract_parallel": par_on}})
    assert cc._tn.infos[0].n_slices >= 8
    xx = x.clone().requires_grad_(True)
    y = cc.batched(xx)
    (y * torch.tensor([0.3, -0.8], device="cuda")).sum().backward()
    res.append((y.detach(), xx.grad.clone()))
err3 = max(float((res[0][0] - res[1][0]).abs().max()), float((res[0][1] - res[1][1]).abs().max()))
# 4. a batch of amplitudes: slices sharded, ONE all-reduce for the whole batch
bits_batch = np.random.RandomState(3).randint(0, 2, size=(5, 16))
b1, b2 = one.amplitudes(bits_batch).cpu(), par.amplitudes(bits_batch).cpu()
err4 = float((b1 - b2).abs().max() / b1.abs().max())
# 5. the C ABI's own multi-GPU entry (tq_dist_create / tq_tn_contract_sharded): a raw ncclComm_t made with ctypes on
#    the NCCL library torch loaded, handed to the library as an opaque pointer
import ctypes
import nvidia.nccl
from tedq_b200 import capi
nccl = ctypes.CDLL(os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2"))
class UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_char * 128)]
uid = UniqueId()
if rank == 0:
    assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
buf = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).cuda()
dist.broadcast(buf, 0)
uid = UniqueId.from_buffer_copy(bytes(buf.cpu().numpy().tobytes()))
comm = ctypes.c_void_p()
nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
assert nccl.ncclCommInitRank(ctypes.byref(comm), dist.get_world_size(), uid, rank) == 0
tqd = capi.Dist(comm.value, rank, dist.get_world_size())
flat0 = torch.zeros((1, 0), device="cuda")
plan, ptrs, strides, out, ws, ws_bytes, _, _keep = one._tn._amplitude_operands(flat0, bits)
out.zero_()
plan.contract_sharded(tqd.handle, ptrs, strides, out.shape[0], out.data_ptr(), ws.data_ptr(), ws_bytes,
                      torch.cuda.current_stream().cuda_stream)
a5 = complex(out.sum().cpu())
err5 = abs(a5 - a1) / abs(a1)
if rank == 0:
    print("C-ABI contract_sharded rel err %.2e %s" % (err5, "OK" if err5 < 1e-5 else "FAIL"))
    print("sharded reverse pass max err %.2e %s" % (err3, "OK" if err3 < 1e-5 else "FAIL"))
    print("amplitudes batch rel err %.2e %s" % (err4, "OK" if err4 < 1e-5 else "FAIL"))
    print("measurement_parallel max err %.2e %s" % (err1, "OK" if err1 < 1e-5 else "FAIL"))
    print("contract_parallel rel err %.2e %s" % (err2, "OK" if err2 < 1e-5 else "FAIL"))
dist.destroy_process_group()
