/*
 * tedq_b200.h — C ABI of the B200-native execution engine behind TeD-Q's
 * `Circuit.compilecircuit(backend="pytorch_b200")`.
 *
 * The reference (jd-opensource/TeD-Q) is pure Python: there is no native
 * interface to mirror, so every entry point below cites the *Python* interface
 * it replaces (file:line under the reference tree).  Plain C types only: no
 * torch types, no C++ types, no exceptions cross this boundary.  The caller
 * owns every device buffer it passes in; kernels are enqueued on the caller's
 * stream and never synchronise it (the *_host entry points are the exception:
 * they take HOST buffers, copy, run and synchronise — they are the "call a
 * user makes" for end-to-end timing).
 *
 * Conventions (reference: tedq/backends/pytorch_backend.py:513-522, :567-577)
 *   - state psi is a C-contiguous [2]*n tensor: qubit q <-> tensor axis q, i.e.
 *     qubit 0 is the MOST significant bit of the flat amplitude index;
 *   - a k-qubit gate matrix is row-major (2^k x 2^k), row = output index,
 *     column = input index, gate qubit 0 = most significant bit of both;
 *   - complex numbers are interleaved (re, im); TQ_C64 = 2 x float,
 *     TQ_C128 = 2 x double.  Parameters / real outputs use the matching real
 *     type (float for TQ_C64, double for TQ_C128).
 *
 * All functions returning int return 0 on success and a negative TQ_E_* code
 * on failure; tq_last_error() gives the message (thread-local).
 */
#ifndef TEDQ_B200_H
#define TEDQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQ_ABI_VERSION 1
#define TQ_MAX_GATE_QUBITS 3  /* reference gate set tops out at CSWAP/Toffoli (qubit.py:838-960) */
#define TQ_MAX_GATE_PARAMS 3  /* Rot (qubit.py:1137-1190) */
#define TQ_MAX_MEAS_QUBITS 64
#define TQ_MAX_OBS_QUBITS 4   /* dense observable matrix up to 16x16 */

enum tq_status {
  TQ_OK = 0,
  TQ_E_INVALID = -1,   /* bad argument / malformed plan input            */
  TQ_E_UNSUPPORTED = -2,
  TQ_E_CUDA = -3,      /* a CUDA runtime call failed                     */
  TQ_E_WORKSPACE = -4, /* workspace too small                            */
  TQ_E_NOMEM = -5
};

enum tq_dtype { TQ_C64 = 0, TQ_C128 = 1 };

/* Gate kinds.  TQ_G_FIXED carries a constant matrix (what the reference does for
 * every gate with no trainable slot: compiled_circuit.py:432-435,
 * pytorch_backend.py:567-577).  The parametrised kinds are evaluated on the
 * device from theta exactly as pytorch_backend.py:866-1188 does on the host. */
enum tq_gate_kind {
  TQ_G_FIXED = 0,
  TQ_G_RX = 1,         /* pytorch_backend.py:866-894   */
  TQ_G_RY = 2,         /* :897-922                     */
  TQ_G_RZ = 3,         /* :925-954                     */
  TQ_G_ROT = 4,        /* :957-984                     */
  TQ_G_PHASESHIFT = 5, /* :987-1013                    */
  TQ_G_CPHASE = 6,     /* :1016-1054                   */
  TQ_G_CRX = 7,        /* :1057-1100                   */
  TQ_G_CRY = 8,        /* :1103-1146                   */
  TQ_G_CRZ = 9         /* :1149-1188                   */
};

typedef struct tq_gate_desc {
  int32_t kind;                           /* enum tq_gate_kind                                  */
  int32_t nq;                             /* 1..TQ_MAX_GATE_QUBITS                              */
  int32_t qubits[4];                      /* qubits[0] = most significant bit of the matrix idx */
  int32_t param_idx[TQ_MAX_GATE_PARAMS];  /* index into the flat parameter vector, -1 = const   */
  int32_t _pad;
  double param_const[TQ_MAX_GATE_PARAMS]; /* value used when param_idx[i] < 0                   */
  int64_t matrix_off;                     /* TQ_G_FIXED: offset (in complex entries) into pool  */
} tq_gate_desc;

enum tq_meas_kind {
  TQ_M_EXPVAL = 0, /* pytorch_backend.py:399-459 */
  TQ_M_PROBS = 1,  /* :461-472                   */
  TQ_M_STATE = 2   /* :495-496                   */
};

#define TQ_MF_ZSTRING 1 /* EXPVAL of a product of PauliZ on `qubits` (any nq) */

typedef struct tq_meas_desc {
  int32_t kind;  /* enum tq_meas_kind */
  int32_t flags; /* TQ_MF_*           */
  int32_t nq;    /* EXPVAL: observable qubits; PROBS: kept qubits (0 = all); STATE: 0 */
  int32_t _pad;
  int32_t qubits[TQ_MAX_MEAS_QUBITS];
  int64_t matrix_off; /* EXPVAL without ZSTRING: 2^nq x 2^nq matrix in pool (nq <= TQ_MAX_OBS_QUBITS) */
} tq_meas_desc;

typedef struct tq_plan_opts {
  int32_t max_local_qubits_fwd; /* 0 = default. log2(amplitudes of one shared-memory tile), forward  */
  int32_t max_local_qubits_bwd; /* 0 = default. same for the adjoint sweep (holds psi and lambda)    */
  int32_t coalesce_bits;        /* -1 = default. low amplitude-index bits always kept tile-local     */
  int32_t threads;              /* 0 = default CTA size                                              */
  int32_t fuse;                 /* -1 = default (1). 1: fuse runs of gates in registers               */
  int32_t structure;            /* 0 = default: automatic.  complex64 circuits of >= 12 qubits whose gates are all
                                   (controlled) one-target blocks or diagonals take the register-group sweeps (2) when
                                   those need no more multiply-adds than the default fusion (layers of one-qubit gates
                                   between sparse entanglers); everything else takes the default sweeps, where every
                                   fused block is a dense complex matrix.
                                   -1: the default sweeps whatever the circuit.
                                   1 (experimental): blocks of real gates (RY, CRY, H, X, CNOT, ...) take real-matrix
                                   paths with half the multiplies, and one-qubit diagonal gates (RZ, PhaseShift, S, T,
                                   Z) stay out of the blocks and are merged, per sweep, into diagonal-layer passes (two
                                   phase tables, signed-sum gradients).  Same results; measured slower on B200.
                                   2: register-group sweeps whenever the circuit qualifies: a thread keeps the 16
                                   amplitudes of four tile bits in registers across a run of blocks; one-qubit runs stay
                                   2x2, X / CNOT blocks at the ends of a run are folded into the load / store addresses.
                                   A circuit that does not qualify silently keeps the default sweeps
                                   (tq_plan_op_stats(what = 4) tells which)                                          */
  int32_t reserved[2];
} tq_plan_opts;

typedef struct tq_plan tq_plan; /* opaque, immutable after creation */

/* ---- library ----------------------------------------------------------- */
int tq_abi_version(void);
const char* tq_last_error(void);

/* ---- plan (replaces CompiledCircuit.__init__ state-vector planning,
 *      compiled_circuit.py:82-202, and the per-call _parser_circuit :418-440) */
int tq_plan_create(const tq_gate_desc* gates, int32_t n_gates, const tq_meas_desc* meas, int32_t n_meas,
                   const double* pool /* complex interleaved */, int64_t pool_complex_len, int32_t n_qubits,
                   int32_t n_params, int32_t dtype, const double* init_state /* nullable, 2^n complex */,
                   const tq_plan_opts* opts /* nullable */, tq_plan** out);
void tq_plan_destroy(tq_plan* plan);

/* introspection (scheduling is part of the contract the tests pin) */
int32_t tq_plan_num_qubits(const tq_plan* plan);
int32_t tq_plan_num_params(const tq_plan* plan);
int32_t tq_plan_num_sweeps(const tq_plan* plan, int32_t backward);
/* ops after gate fusion (runs of gates inside one qubit / one qubit pair become one dense block) */
int32_t tq_plan_num_blocks(const tq_plan* plan);
/* ops emitted into the sweeps of one direction: what = 0 all, 1 diagonal-layer ops (runs of one-qubit diagonal gates
 * merged into one phase-table pass), 2 the gates inside them, 3 ops on the real-matrix paths, 4 register groups
 * (structure = 2; 0 when the plan kept the default sweeps) */
int32_t tq_plan_op_stats(const tq_plan* plan, int32_t backward, int32_t what);
/* local amplitude-index bit positions of sweep s; returns count, writes up to cap entries */
int32_t tq_plan_sweep_bits(const tq_plan* plan, int32_t backward, int32_t s, int32_t* bits, int32_t cap);
int32_t tq_plan_sweep_num_gates(const tq_plan* plan, int32_t backward, int32_t s);
/* number of REAL scalars one parameter set produces (all measurements, stacked) */
int64_t tq_plan_out_reals(const tq_plan* plan);
/* algorithmic HBM bytes one forward / backward evaluation moves (sweeps x read+write of psi [and lambda]) */
int64_t tq_plan_hbm_bytes(const tq_plan* plan, int32_t backward);
int64_t tq_plan_launches(const tq_plan* plan, int32_t backward);
/* algorithmic real flops of one forward / backward evaluation on the fused-block schedule (complex MAC = 8 flops);
 * the sweeps of a shared-memory-resident state are bound by the FP32 / FP64 pipe, not by HBM */
double tq_plan_flops(const tq_plan* plan, int32_t backward);

/* ---- execution (replaces PyTorchBackend.execute state-vector branch,
 *      pytorch_backend.py:358-391 + get_measurement_results :393-498;
 *      backward replaces autograd through that loop) -------- */
size_t tq_workspace_bytes(const tq_plan* plan, int64_t batch, int32_t with_backward);

/* params: device [batch, n_params] real. out: device [batch, out_reals] real.
 * workspace: device, >= tq_workspace_bytes(plan, batch, with_backward); after a
 * forward with with_backward != 0 it holds what tq_backward needs. */
int tq_forward(const tq_plan* plan, const void* params, int64_t batch, void* out, void* workspace,
               size_t workspace_bytes, int32_t with_backward, void* cuda_stream);

/* grad_out: device [batch, out_reals] real (torch convention: for a complex
 * STATE output the cotangent is (dL/dRe, dL/dIm) interleaved).
 * grad_params: device [batch, n_params] real, overwritten. */
int tq_backward(const tq_plan* plan, const void* params, int64_t batch, const void* grad_out,
                void* grad_params, void* workspace, size_t workspace_bytes, void* cuda_stream);

/* final state of the last tq_forward (device pointer inside workspace, [batch, 2^n] complex), or NULL
 * when the state never left shared memory */
void* tq_workspace_state(const tq_plan* plan, void* workspace, int64_t batch);

/* HOST-buffer entry point: H2D(params[, grad_out]) -> forward [-> backward] -> D2H(out[, grad_params]),
 * synchronises.  grad_out / grad_params may both be NULL (forward only).  Device scratch is owned by the
 * plan and reused between calls (this one call is therefore NOT re-entrant on one plan). */
int tq_execute_host(tq_plan* plan, const void* params, int64_t batch, void* out, const void* grad_out,
                    void* grad_params);

/* ---- state-vector plan integers (bit-exact mirror of compiled_circuit.py:126-202) */
/* gate_pos[k] and perm[n_qubits] for one gate; returns 0 */
int tq_sv_axes_perm(int32_t n_qubits, const int32_t* qubits, int32_t nq, int32_t* gate_pos, int32_t* perm);

/* ---- tensor-network side (replaces gen_tensor_networks index maps,
 *      tensor_network.py:850-1099, and the third-party tree.contract call sites
 *      pytorch_backend.py:276,:339 / oe_wrapper.py:59-65) -------------------- */

#define TQ_TN_MAX_RANK 32

/* unicode code point of symbol i (get_symbol, tensor_network.py:1109-1127) */
int32_t tq_tn_symbol(int32_t i);

/* Index maps of ONE measurement's network (host only, no GPU needed).
 *   gate_nq[g], gate_qubits[g*4 + i]: the circuit's gates in order.
 *   meas_kind: TQ_M_*; obs_nq[j] / obs_qubits[j*4 + i]: the observable tensors of an EXPVAL (one entry per
 *   observable of a list observable); kept_qubits / n_kept: PROBS (n_kept < 0 means qubits=None).
 * Writes tensor t's index ids, slow -> fast (the reference's list order), at
 * tensor_idx[tensor_off[t] .. tensor_off[t+1]) and the open indices to out_idx.
 * Returns the number of tensors, or a negative tq_status. */
int32_t tq_tn_index_map(int32_t n_qubits, const int32_t* gate_nq, const int32_t* gate_qubits, int32_t n_gates,
                        int32_t meas_kind, const int32_t* obs_nq, const int32_t* obs_qubits, int32_t n_obs,
                        const int32_t* kept_qubits, int32_t n_kept, int32_t* tensor_off, int32_t* tensor_idx,
                        int32_t tensor_cap, int32_t idx_cap, int32_t* out_idx, int32_t* n_out);

/* One lowered pairwise-contraction step  C[b, m, n] = sum_k A[b, m, k] * B[b, k, n].
 * Every extent is 2: tensors are addressed by bit strings, a mode permutation is a bit permutation.
 * lhs_bits = physical bit positions inside A of [k..., m..., b...] (each group fast -> slow),
 * rhs_bits = the same for B with n instead of m.  The result is dense, laid out [b | m | n] slow -> fast;
 * out_idx[j] = index id of result bit j (fast -> slow). */
typedef struct tq_tn_step {
  int32_t lhs, rhs;               /* ssa ids: inputs are 0..n_in-1, step s produces n_in+s */
  int32_t n_k, n_m, n_n, n_b;     /* log2 extents                                           */
  int8_t lhs_bits[TQ_TN_MAX_RANK];
  int8_t rhs_bits[TQ_TN_MAX_RANK];
  int32_t out_idx[TQ_TN_MAX_RANK];
} tq_tn_step;

/* Planner, host only: the dynamic programme of the subtree reconfiguration (ted-q_b200/planner.py: reconfigure; the
 * search the reference delegates to cotengra / jdtensorpath, compiled_circuit.py:340-393).  A subtree has n_leaves
 * (2..12) leaf tensors over n_idx distinct indices: leaf_open[l][x] != 0 when index x is an open leg of leaf l,
 * inside[l][x] = how many input tensors below leaf l carry x, count[x] = how many tensors of the whole network (plus
 * the output) carry x.  Finds for every subset S of the leaves the cheapest way to contract it pairwise
 * (cost of a pair = 2^|union of open legs|, or with time_model = {tensor-core flop/s, bytes/s, s per step, FP32-GEMM
 * flop/s or 0, per-element-kernel flop/s: FIVE values} the estimated step time of planner.step_time_model); *best_full = cost of the whole subtree, split[S] = the first part of S's best
 * split (the part that contains S's lowest leaf), for every S with at least two leaves.  Bit-identical to the
 * Python mirror planner._subtree_dp_py.  Returns 0 or a negative tq_status. */
int32_t tq_tn_subtree_order(int32_t n_leaves, int32_t n_idx, const int32_t* leaf_open, const int32_t* inside,
                            const int32_t* count, const double* time_model, double* best_full, int32_t* split);

/* Planner, host only: one randomised greedy pass (ted-q_b200/planner.py: _greedy_once, whose path it reproduces bit for
 * bit — same operations, same libm calls).  Inputs i = 0 .. n_inputs-1 carry the indices idx[idx_off[i] .. idx_off[i+1])
 * (renumbered 0 .. n_idx-1), keep[x] != 0 marks the network's output indices.  Candidate pairs are scored
 * log2-compressed |out| - alpha (|a| + |b|) minus temperature x Gumbel noise and contracted best first; init_pairs lists
 * the pairs that share an index in the order the caller wants them scored (2 ids per pair), u[] supplies one uniform
 * random number per scored pair when temperature > 0.  Writes the ssa path (2 (n_inputs - 1) ids; the k-th contraction
 * creates id n_inputs + k) and the count of random numbers consumed.  TQ_E_WORKSPACE: u was too short. */
int32_t tq_tn_greedy_path(int32_t n_inputs, int32_t n_idx, const int32_t* idx_off, const int32_t* idx, const int32_t* keep,
                          int32_t n_init, const int32_t* init_pairs, const double* u, int64_t n_u, double alpha,
                          double temperature, int32_t* path, int64_t* n_u_used);

/* Lower an ssa path over n_in input tensors into steps (host only).  sliced[] index ids become per-slice base
 * offsets of the inputs that carry them: slice_tensor[i], slice_ord[i], slice_bit[i] (count returned in
 * *n_slice_entries, capacity slice_cap).  final_perm[j] = bit of the last tensor that becomes output bit j
 * (output order = out_idx listed slow -> fast).  Returns the number of steps or a negative tq_status. */
int32_t tq_tn_lower(const int32_t* tensor_off, const int32_t* tensor_idx, int32_t n_in, const int32_t* out_idx,
                    int32_t n_out, const int32_t* ssa_path /* 2 per step */, int32_t n_steps,
                    const int32_t* sliced, int32_t n_sliced, tq_tn_step* steps, int32_t* slice_tensor,
                    int32_t* slice_ord, int32_t* slice_bit, int32_t slice_cap, int32_t* n_slice_entries,
                    int32_t* final_perm);

typedef struct tq_tn_plan tq_tn_plan;

/* Executable contraction plan.  input_batched[t] != 0: input t differs per parameter set (it carries a
 * leading batch dimension in tq_tn_contract). */
int tq_tn_plan_create(const int32_t* tensor_off, const int32_t* tensor_idx, int32_t n_in, const int32_t* out_idx,
                      int32_t n_out, const int32_t* ssa_path, int32_t n_steps, const int32_t* sliced,
                      int32_t n_sliced, const int32_t* input_batched, int32_t dtype, tq_tn_plan** out);
void tq_tn_plan_destroy(tq_tn_plan* plan);
int32_t tq_tn_plan_num_steps(const tq_tn_plan* plan);
int64_t tq_tn_plan_num_slices(const tq_tn_plan* plan);
double tq_tn_plan_flops(const tq_tn_plan* plan);  /* sum over steps of 8*2^(k+m+n+b), ONE slice, one set */
int32_t tq_tn_plan_width(const tq_tn_plan* plan); /* log2 of the largest tensor of one slice           */
int32_t tq_tn_plan_get_step(const tq_tn_plan* plan, int32_t s, tq_tn_step* out);
size_t tq_tn_workspace_bytes(const tq_tn_plan* plan, int64_t batch);

/* Execution options (set before tq_tn_workspace_bytes / tq_tn_contract).
 *   TQ_TN_OPT_TENSOR_CORE (default 1): complex64 steps with >= 128 x 16 free extents and
 *     k + m + n + b >= TQ_TN_OPT_TC_MIN_LOG2 (default 20) run on tcgen05 tensor cores as a 4M real GEMM with
 *     error-compensated split-TF32 (x = hi + lo; hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM);
 *     0 keeps every step on the fp32 FMA kernels (used by the parity tests to compare the two paths).
 *   TQ_TN_OPT_TC_CHUNK (default 32): complex k accumulated inside the tensor core between drains.  tcgen05
 *     accumulates with round-toward-zero (a bias linear in K); partial sums are therefore drained every
 *     `chunk` complex k and added to fp32 registers with round-to-nearest.
 *   TQ_TN_OPT_FUSE_SMALL (default 1): steps with k + m + n + b <= 18 (batched over parameter sets) or <= 14
 *     (shared) are grouped into fused runs — ONE launch walks a dependency-closed set of small steps level by
 *     level, one CTA per parameter set, tiny steps one per warp; 0 launches every step on its own.
 *   TQ_TN_OPT_TC_SPLITK (default 1): a tensor-core step whose output tiles cover at most half of the SMs is cut
 *     along K (>= 128 complex k per part) so that every SM has work; the parts write partial sums that are added
 *     in a fixed order (deterministic).
 *   TQ_TN_OPT_TC_GATHER (default 0, experimental): 1 = the row operand of a tensor-core step whose tiles are read
 *     once or twice (<= 2 column tiles) is gathered, split and swizzled into shared memory by the GEMM kernel itself,
 *     straight from the tensor (no operand image in HBM for it).  Bit-identical results; measured slower than the
 *     image path on B200 so far ("Experiments that did not pay").
 *   TQ_TN_OPT_TC_FUSE_PACK (default 1): a tensor-core step whose result is read by exactly one step, itself a
 *     tensor-core step, writes that step's operand image (hi / lo TF32 planes, permuted, swizzled) straight from its
 *     epilogue: the intermediate is never stored in its plain layout, the consumer runs no pack pass.  0 = every
 *     tensor-core step packs its operands from plain tensors (bit-identical results; parity tests compare both).
 *   TQ_TN_OPT_CHAIN (default 1): runs of consecutive gate-like apply steps on the same large tensor (each contracts
 *     k <= 4 of its indices with a small operand and puts k new ones in their place: every step of a
 *     state-vector-like plan) execute as ONE launch of k_tn_chain — a shared-memory tile per CTA, the tensors between
 *     the steps are never materialised.  0 = one k_tn_apply launch per step. */
enum tq_tn_option { TQ_TN_OPT_TENSOR_CORE = 0, TQ_TN_OPT_TC_MIN_LOG2 = 1, TQ_TN_OPT_TC_CHUNK = 2,
                    TQ_TN_OPT_FUSE_SMALL = 3, TQ_TN_OPT_TC_SPLITK = 4, TQ_TN_OPT_TC_GATHER = 5,
                    TQ_TN_OPT_TC_FUSE_PACK = 6, TQ_TN_OPT_CHAIN = 7 };
int tq_tn_plan_set_option(tq_tn_plan* plan, int32_t option, int32_t value);
/* kernel that runs step s: 0 = one thread per output element, 1 = tiled fp32 FMA GEMM,
 * 2 = tcgen05 split-TF32 GEMM over packed operand images, 3 = split-K reduction (<= 64 outputs, K >= 4096),
 * 4 = member of a fused run of small steps, 5 = apply kernel (one gate-sized operand, one large operand),
 * 6 = gradient seed, 7 = member of an apply-chain run (k_tn_chain) */
int32_t tq_tn_plan_step_kernel(const tq_tn_plan* plan, int32_t s);
/* fused pack: the step whose operand image step s writes from its own epilogue, -1 when s stores a plain tensor
 * (a split-K launch of s still falls back to plain + pack at run time) */
int32_t tq_tn_plan_step_fuse_to(const tq_tn_plan* plan, int32_t s);
/* how the epilogue of a fused-pack producer reaches whole 64-byte image rows (0: not fused): 1 = the consumer's three
 * low k bits are accumulator ROW bits of the producer (8 lanes x 8 bytes), 2 = the lowest k bit is a column bit and the
 * next two are row bits (4 lanes x 16 bytes), 3 = all three are column bits (one thread, 16-byte stores) */
int32_t tq_tn_plan_step_fuse_mode(const tq_tn_plan* plan, int32_t s);
/* bit 0: step s repeats for every slice (it depends on a sliced index); bit 1: it carries the parameter-set
 * batch dimension.  Steps with neither bit run once per call, outside the slice loop. */
int32_t tq_tn_plan_step_flags(const tq_tn_plan* plan, int32_t s);

/* Contract slices [slice_begin, slice_end) and ACCUMULATE their sum into out (device,
 * [batch or 1][2^n_out] complex; the caller zeroes it).  inputs[t] = device pointer of input tensor t
 * (C order, complex); input_strides[t] = complex entries between consecutive parameter sets of input t
 * (0 for inputs shared by every set).  Multi-GPU: each rank takes a slice range and the
 * caller all-reduces out with ONE ncclAllReduce (the reference's analogue is jdtensorpath's RPC slice sum,
 * examples/qubit_rpc.py:110-126). */
int tq_tn_contract(const tq_tn_plan* plan, const void* const* inputs, const int64_t* input_strides, int64_t batch,
                   int64_t slice_begin, int64_t slice_end, void* out, void* workspace, size_t workspace_bytes,
                   void* cuda_stream);

/* Two-stage form of tq_tn_contract, for callers that contract the same network many times (a batch of amplitudes):
 * tq_tn_contract_prepare runs the once-per-call part (the steps that depend on no sliced index, the pinned operand
 * images) into `workspace`; tq_tn_contract_slices then runs only the slice loop on that workspace with the same
 * inputs.  With two workspaces and two streams the (latency-bound) preparation of the next contraction overlaps the
 * (throughput-bound) slices of the current one; the caller orders the two calls with events.  slice_begin of
 * _prepare: the first slice the following _slices call will run (it selects the sliced entries of pinned images). */
int tq_tn_contract_prepare(const tq_tn_plan* plan, const void* const* inputs, const int64_t* input_strides,
                           int64_t batch, int64_t slice_begin, void* workspace, size_t workspace_bytes,
                           void* cuda_stream);
int tq_tn_contract_slices(const tq_tn_plan* plan, const void* const* inputs, const int64_t* input_strides,
                          int64_t batch, int64_t slice_begin, int64_t slice_end, void* out, void* workspace,
                          size_t workspace_bytes, void* cuda_stream);

/* ---- multi-GPU (reference analogue: jdtensorpath's RPC slice workers, examples/qubit_rpc.py:110-126,
 *      tedq/distributed_worker/rpc_workers.py:55-59) ----------------------------------------------------------
 * One process per GPU.  The host creates the NCCL communicator (ncclCommInitRank, or the one its framework owns) and
 * hands it over as an opaque pointer; the library resolves ncclAllReduce from the NCCL already loaded in the process
 * (dlopen of libnccl.so.2: no link-time dependency).  tq_tn_contract_sharded contracts THIS rank's contiguous range
 * of slices (tq_dist_slice_range: ceil(n / world) per rank) into `out` (zeroed by the caller) and combines the
 * partial sums of all ranks with ONE ncclAllReduce(sum) enqueued on the same stream; every rank ends up with the
 * full result.  An unsliced plan is contracted by rank 0 only (the others contribute zeros). */
typedef struct tq_dist tq_dist;
int tq_dist_create(void* nccl_comm /* ncclComm_t */, int32_t rank, int32_t world, tq_dist** out);
void tq_dist_destroy(tq_dist* dist);
int tq_dist_slice_range(int64_t n_slices, int32_t rank, int32_t world, int64_t* begin, int64_t* end);
/* in-place sum over ranks of `count` reals (dtype: TQ_C64 -> float, TQ_C128 -> double) on the stream */
int tq_dist_allreduce(const tq_dist* dist, void* buf, int64_t count, int32_t dtype, void* cuda_stream);
int tq_tn_contract_sharded(const tq_tn_plan* plan, const tq_dist* dist, const void* const* inputs,
                           const int64_t* input_strides, int64_t batch, void* out, void* workspace,
                           size_t workspace_bytes, void* cuda_stream);

/* ---- reverse mode through the contraction tree (the reference differentiates through tree.contract with torch's
 *      tape, pytorch_backend.py:276/:339 under back_prop) ------------------------------------------------------
 * tq_tn_plan_enable_backward appends, for every forward step C = A x B on the way to an input that needs a
 * gradient, the two contractions  g_A = g_C x B  and  g_B = A x g_C  of the CONJUGATED gradients g = conj(dL/dT)
 * (so that no operand has to be conjugated), plus the seed g_out = conj(grad_out); the schedule, the fused runs and
 * the arena layout are rebuilt for the combined list, forward intermediates that the reverse pass reads stay
 * alive.  Usage, per slice s: tq_tn_contract(..., slices [s, s + 1), workspace W) then
 * tq_tn_backward(..., s, grad_out, the SAME workspace W); gradients of the slices add up (the contraction is a sum
 * over slices), a sliced index of an input is fixed per slice and absent from its gradient tensor (bit -1 in
 * tq_tn_grad_info).  Gradients stay in the workspace:
 *   tq_tn_workspace_layout -> byte offsets of the shared / per-set arenas and the per-set stride (complex entries)
 *   tq_tn_grad_info(t)     -> element offset, arena (-1 shared, -2 per set) and, for the i-th index of input t's
 *                             own list, its bit inside the gradient tensor (entries are conj(dL/dT)).
 * tq_tn_param_grads (circuit networks) applies the chain rule to the flat parameters: off_g / off_a are device
 * int32 [n_gates][16] tables of per-set-arena element offsets of the gradient entries of G / G^dagger (-1: none);
 * grad_params [batch][n_params] is accumulated (+=), the caller zeroes it. */
int tq_tn_plan_enable_backward(tq_tn_plan* plan, const int32_t* input_needs_grad);
int tq_tn_backward(const tq_tn_plan* plan, const void* const* inputs, const int64_t* input_strides, int64_t batch,
                   int64_t slice, const void* grad_out, void* workspace, size_t workspace_bytes, void* cuda_stream);
int tq_tn_grad_info(const tq_tn_plan* plan, int32_t t, int64_t* offset, int32_t* space, int32_t* bits);
int tq_tn_workspace_layout(const tq_tn_plan* plan, int64_t* shared_off, int64_t* perset_off, int64_t* set_stride);
int tq_tn_param_grads(const tq_plan* plan, const void* params, int64_t batch, const void* arena, int64_t set_stride,
                      const int32_t* off_g, const int32_t* off_a, void* grad_params, void* cuda_stream);

/* Profiling twin of tq_tn_contract for ONE slice: same work, plus CUDA events around every step.
 * step_ms has 2 * n_steps + 2 entries: step_ms[2*s] = milliseconds of step s, step_ms[2*s+1] = the part spent
 * packing operand images (tensor-core steps only); step_ms[2*n_steps] = once-per-call packing of the
 * slice-invariant operand images that per-slice steps reuse.  Synchronises the stream.  Feeds bench.py's per-step
 * roofline table. */
int tq_tn_profile(const tq_tn_plan* plan, const void* const* inputs, const int64_t* input_strides, int64_t batch,
                  int64_t slice, void* out, void* workspace, size_t workspace_bytes, void* cuda_stream,
                  float* step_ms);

/* Kernel launches enqueued by tq_tn_contract / tq_tn_backward / tq_tn_profile since the library was loaded
 * (monotonic; a caller diffs it around a region: bench.py's gpu_launches).  The state-vector side reports its
 * launches per call through tq_plan_launches. */
int64_t tq_tn_launch_count(void);

/* Operand tensors of a circuit's network on the device (replaces _parse_circuit_cotengra + the arrays
 * assembly, compiled_circuit.py:442-467, pytorch_backend.py:311-336, :524-546): for every gate of `plan`
 * writes G (row-major [out..., in...]) and G^dagger, per parameter set.
 *   gate_mats / adj_mats: device [batch][total] complex, gate g at offset tq_tn_gate_offset(plan, g). */
int64_t tq_tn_gate_offset(const tq_plan* plan, int32_t gate);   /* gate == n_gates -> total entries */
int tq_tn_operands(const tq_plan* plan, const void* params, int64_t batch, void* gate_mats, void* adj_mats,
                   void* cuda_stream);

/* Operands of a structure-aware ("simplified", tn_simplify=True) network: diagonal gates keep only their diagonal,
 * controlled gates only the entries with equal control bits on both sides — sub-sets of the tensors written by
 * tq_tn_operands.  dst[set][i] = gate_mats[set][idx[i]] for idx[i] >= 0, adj_mats[set][-idx[i]-1] otherwise
 * (idx: device int32[n]; src_stride = entries per parameter set of both source buffers; dst: [batch][n]).
 * The reference's own simplifier (tensor_network.py:90-129) is the non-working counterpart. */
int tq_tn_gather(const void* gate_mats, const void* adj_mats, int64_t src_stride, const int32_t* idx, int64_t n,
                 void* dst, int64_t batch, int32_t dtype, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* TEDQ_B200_H */
