//! Bodies that replace the CPU calls in src/python.rs for the hot path: to_csc / to_csr (:27-53),
//! neighbor_sampling_homogenous (:187-271), neighbor_sampling_heterogenous (:273-395) and random_walk (:583-608).
//! Signatures, argument extraction (`SamplerType`, `FilterType`, `GraphSize`) and return layouts of python.rs are
//! untouched: each `#[pyfunction]` keeps its header and calls the `*_cuda` function below instead of `crate::algo::…`.
//! Everything stateful lives behind the C ABI: a `Graph` (tchgeo_graph_t: borrowed CSC arrays + the derived int32 replica
//! / weight records, built once per set of tensors) and a plan (tchgeo_plan_t: deep-copied arguments, launch plan,
//! pinned length table).  Work is issued on torch's CURRENT stream (csrc/torch_stream_shim.cpp), never on the legacy one.
//! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no cargo/rustc); the Python mirror (tch_geometric/ops.py) makes exactly
//! these calls through ctypes and tests/cpp/abi_harness.cpp makes them from C++.
use std::collections::HashMap;
use std::os::raw::c_void;
use std::sync::Mutex;

use lazy_static::lazy_static;
use rand::RngCore;
use tch::{Device, Kind, Tensor};

use crate::cuda_ffi as ffi;
use crate::utils::{random, TensorConversionError, TensorResult};

/// CUDA counterpart of try_tensor_to_slice (src/utils/tensor.rs:50-59): device + dtype (+ contiguity) check, raw pointer.
fn cuda_ptr<T>(t: &Tensor, kind: Kind) -> TensorResult<*const T> {
    if !matches!(t.device(), Device::Cuda(_)) { return Err(TensorConversionError::InvalidDevice(Device::Cuda(0))); }
    if t.kind() != kind { return Err(TensorConversionError::InvalidDType(kind, t.kind())); }
    if !t.is_contiguous() { return Err(TensorConversionError::Unknown("tensor must be contiguous".into())); }
    Ok(t.data_ptr() as *const T)
}

fn device_index(t: &Tensor) -> i32 { match t.device() { Device::Cuda(i) => i as i32, _ => 0 } }
fn stream_of(t: &Tensor) -> ffi::tchgeo_stream { unsafe { ffi::tchgeo_torch_current_stream(device_index(t)) } }

/// Owns a tchgeo_graph_t and shallow clones of the tensors it borrows (so their storage outlives the handle).
pub struct Graph { handle: *mut ffi::tchgeo_graph_t, _keep: Vec<Tensor> }
unsafe impl Send for Graph {}
impl Drop for Graph { fn drop(&mut self) { unsafe { ffi::tchgeo_graph_destroy(self.handle) } } }

lazy_static! {
    /// one handle per set of graph tensors, found again by storage address + length (the reference API is stateless:
    /// tensors arrive on every call).  Same policy as tch_geometric/ops.py::_GraphCache.
    static ref GRAPHS: Mutex<HashMap<Vec<(usize, i64)>, std::sync::Arc<Graph>>> = Mutex::new(HashMap::new());
}

fn graph_for(ptrs: &[&Tensor], indices: &[&Tensor], weights: Option<&[&Tensor]>) -> TensorResult<std::sync::Arc<Graph>> {
    let mut key: Vec<(usize, i64)> = ptrs.iter().chain(indices.iter()).map(|t| (t.data_ptr() as usize, t.numel() as i64)).collect();
    if let Some(w) = weights { key.extend(w.iter().map(|t| (t.data_ptr() as usize, t.numel() as i64))); }
    let mut cache = GRAPHS.lock().unwrap();
    if let Some(g) = cache.get(&key) { return Ok(g.clone()); }
    let p: Vec<*const i64> = ptrs.iter().map(|t| cuda_ptr::<i64>(t, Kind::Int64)).collect::<Result<_, _>>()?;
    let i: Vec<*const i64> = indices.iter().map(|t| cuda_ptr::<i64>(t, Kind::Int64)).collect::<Result<_, _>>()?;
    let num_major: Vec<i64> = ptrs.iter().map(|t| t.numel() as i64 - 1).collect();
    let nnz: Vec<i64> = indices.iter().map(|t| t.numel() as i64).collect();
    let mut handle = std::ptr::null_mut();
    ffi::check(unsafe { ffi::tchgeo_graph_create(p.len() as i32, p.as_ptr(), num_major.as_ptr(), i.as_ptr(), nnz.as_ptr(), &mut handle) })?;
    let mut keep: Vec<Tensor> = ptrs.iter().chain(indices.iter()).map(|t| t.shallow_clone()).collect();
    if let Some(w) = weights {
        // EdgeAttr::get slices weights with the column's range and panics when they are short (src/data/graph.rs:103-120)
        for (wt, it) in w.iter().zip(indices.iter()) { assert!(wt.numel() >= it.numel(), "weights shorter than row_indices"); }
        let wp: Vec<*const f64> = w.iter().map(|t| cuda_ptr::<f64>(t, Kind::Double)).collect::<Result<_, _>>()?;
        ffi::check(unsafe { ffi::tchgeo_graph_set_weights(handle, wp.as_ptr()) })?;
        keep.extend(w.iter().map(|t| t.shallow_clone()));
    }
    let g = std::sync::Arc::new(Graph { handle, _keep: keep });
    cache.insert(key, g.clone());
    Ok(g)
}

/// replaces python.rs:27-39 (to_csc, csc = true) and :41-53 (to_csr, csc = false): SparseGraphStorage::try_from
/// (src/data/storage.rs:103-127) -> (ptrs, indices, perm), all CUDA i64.
pub fn to_csx_cuda(row_col: &Tensor, size: (i64, i64), csc: bool) -> TensorResult<(Tensor, Tensor, Tensor)> {
    let dev = row_col.device();
    let base = cuda_ptr::<i64>(row_col, Kind::Int64)?;          // [2, E] row-major: rows = src, cols = dst
    let e = row_col.size()[1];
    let n_major = if csc { size.1 } else { size.0 };
    let (ptrs, indices, perm) = (Tensor::empty(&[n_major + 1], (Kind::Int64, dev)), Tensor::empty(&[e], (Kind::Int64, dev)),
                                 Tensor::empty(&[e], (Kind::Int64, dev)));
    let ws_bytes = unsafe { ffi::tchgeo_coo_to_csx_workspace_bytes(e, size.0, size.1) };
    let ws = Tensor::empty(&[ws_bytes.max(1) as i64], (Kind::Uint8, dev));
    ffi::check(unsafe { ffi::tchgeo_coo_to_csx(base, base.add(e as usize), e, size.0, size.1, csc as i32,
        ptrs.data_ptr() as *mut i64, indices.data_ptr() as *mut i64, perm.data_ptr() as *mut i64,
        ws.data_ptr(), ws_bytes, stream_of(row_col)) })?;
    Ok((ptrs, indices, perm))
}

/// Everything one sampling call hands to the library: outputs at worst-case capacity (tchgeo_neighbor_sampling_capacity),
/// workspace, and the plan that borrows them.
struct Call { plan: *mut ffi::tchgeo_plan_t, samples: Vec<Tensor>, rows: Vec<Tensor>, cols: Vec<Tensor>, eidx: Vec<Tensor>,
              _states: Vec<Tensor>, _ws: Tensor, _graph: std::sync::Arc<Graph> }
impl Drop for Call { fn drop(&mut self) { unsafe { ffi::tchgeo_plan_destroy(self.plan) } } }

#[allow(clippy::too_many_arguments)]
fn make_call(graph: std::sync::Arc<Graph>, dev: Device, rel_src: &[i32], rel_dst: &[i32], fanouts: &[i64], rel_active: &[u8],
             num_hops: i32, sampler_kind: i32, inputs: &[Option<&Tensor>], filter: Option<(i32, bool, (i64, i64), Vec<*const i64>, Vec<Option<&Tensor>>)>)
    -> TensorResult<Call> {
    let (t, r) = (inputs.len(), rel_src.len());
    let seeds: Vec<i64> = inputs.iter().map(|x| x.map_or(0, |x| x.numel() as i64)).collect();
    let in_ptrs: Vec<*const i64> = inputs.iter().map(|x| x.map_or(Ok(std::ptr::null()), |x| cuda_ptr::<i64>(x, Kind::Int64))).collect::<Result<_, _>>()?;
    let mut a: ffi::tchgeo_sampling_args = unsafe { std::mem::zeroed() };
    a.num_node_types = t as i32; a.num_rels = r as i32; a.num_hops = num_hops; a.sampler_kind = sampler_kind;
    a.rel_src = rel_src.as_ptr(); a.rel_dst = rel_dst.as_ptr(); a.graph = graph.handle;
    a.fanouts = fanouts.as_ptr(); a.rel_active = rel_active.as_ptr();
    a.num_batches = 1; a.inputs = in_ptrs.as_ptr(); a.seeds_per_batch = seeds.as_ptr();
    let (mut cap_n, mut cap_e) = (vec![0i64; t], vec![0i64; r]);
    ffi::check(unsafe { ffi::tchgeo_neighbor_sampling_capacity(&a, cap_n.as_mut_ptr(), cap_e.as_mut_ptr()) })?;
    let i64s = |n: i64| Tensor::empty(&[n], (Kind::Int64, dev));
    let samples: Vec<Tensor> = cap_n.iter().map(|&n| i64s(n)).collect();
    let (rows, cols, eidx): (Vec<Tensor>, Vec<Tensor>, Vec<Tensor>) =
        (cap_e.iter().map(|&n| i64s(n)).collect(), cap_e.iter().map(|&n| i64s(n)).collect(), cap_e.iter().map(|&n| i64s(n)).collect());
    let mp = |v: &Vec<Tensor>| v.iter().map(|x| x.data_ptr() as *mut i64).collect::<Vec<_>>();
    let (ps, pr, pc, pe) = (mp(&samples), mp(&rows), mp(&cols), mp(&eidx));
    a.samples = ps.as_ptr(); a.samples_stride = cap_n.as_ptr();
    a.rows = pr.as_ptr(); a.cols = pc.as_ptr(); a.edge_index = pe.as_ptr(); a.edges_stride = cap_e.as_ptr();
    let mut states: Vec<Tensor> = vec![];
    let (st_ptrs, is_ptrs): (Vec<*mut i64>, Vec<*const i64>);
    if let Some((mode, forward, window, ts, inputs_state)) = &filter {
        // python.rs:219-249; ABI mode = reference mode + 1
        a.filter_mode = *mode + 1; a.filter_forward = *forward as i32;
        a.filter_window_lo = window.0; a.filter_window_hi = window.1;
        a.timestamps = ts.as_ptr();
        states = cap_n.iter().map(|&n| i64s(n)).collect();
        st_ptrs = mp(&states);
        is_ptrs = inputs_state.iter().map(|x| x.map_or(Ok(std::ptr::null()), |x| cuda_ptr::<i64>(x, Kind::Int64))).collect::<Result<_, _>>()?;
        a.states = st_ptrs.as_ptr(); a.inputs_state = is_ptrs.as_ptr();
    }
    let stream = unsafe { ffi::tchgeo_torch_current_stream(match dev { Device::Cuda(i) => i as i32, _ => 0 }) };
    a.stream = stream;
    a.workspace_bytes = unsafe { ffi::tchgeo_neighbor_sampling_workspace_bytes(&a) };   // builds the graph's derived arrays
    let ws = Tensor::empty(&[a.workspace_bytes.max(1) as i64], (Kind::Uint8, dev));
    a.workspace = ws.data_ptr() as *mut c_void;
    let mut plan = std::ptr::null_mut();
    ffi::check(unsafe { ffi::tchgeo_plan_create(&a, &mut plan) })?;                     // deep-copies every host table above
    Ok(Call { plan, samples, rows, cols, eidx, _states: states, _ws: ws, _graph: graph })
}

fn run(call: &Call, t: usize, r: usize, h: usize, stream: ffi::tchgeo_stream)
    -> TensorResult<(Vec<i64>, Vec<i64>, Vec<Vec<(i64, i64, i64)>>)> {
    let seed = random::rng_get().next_u64();            // the global RNG still forks one child per call (random.rs:19-23)
    ffi::check(unsafe { ffi::tchgeo_plan_enqueue(call.plan, seed, 0, stream) })?;
    ffi::check(unsafe { ffi::tchgeo_plan_collect(call.plan) })?;
    let (mut ns, mut ne, mut lo) = (std::ptr::null(), std::ptr::null(), std::ptr::null());
    ffi::check(unsafe { ffi::tchgeo_plan_results(call.plan, &mut ns, &mut ne, &mut lo, std::ptr::null_mut()) })?;
    let ns = unsafe { std::slice::from_raw_parts(ns, t) }.to_vec();
    let ne = unsafe { std::slice::from_raw_parts(ne, r) }.to_vec();
    let lo = unsafe { std::slice::from_raw_parts(lo, r * h.max(1) * 3) };
    let offsets = (0..r).map(|i| (0..h).map(|k| { let o = (i * h + k) * 3; (lo[o], lo[o + 1], lo[o + 2]) }).collect()).collect();
    Ok((ns, ne, offsets))
}

/// replaces python.rs:202-262 (rng_get, slices, CscGraph::new, the 18-way dispatch, the algo call, Vec -> Tensor copies).
/// sampler_kind / weights come from `SamplerType` (:107-135), filter from `FilterType` (:137-168).
pub fn neighbor_sampling_homogenous_cuda(col_ptrs: &Tensor, row_indices: &Tensor, inputs: &Tensor, num_neighbors: &[usize],
                                         sampler_kind: i32, weights: Option<&Tensor>,
                                         filter: Option<(i32, bool, (i64, i64), &Tensor, &Tensor)>)
    -> TensorResult<(Tensor, Tensor, Tensor, Tensor, Vec<(i64, i64, i64)>)> {
    let graph = graph_for(&[col_ptrs], &[row_indices], weights.map(|w| vec![w]).as_deref())?;
    let fan: Vec<i64> = num_neighbors.iter().map(|&k| k as i64).collect();
    let flt = match filter {
        Some((mode, fwd, win, ts, st)) => Some((mode, fwd, win, vec![cuda_ptr::<i64>(ts, Kind::Int64)?], vec![Some(st)])),
        None => None,
    };
    let call = make_call(graph, col_ptrs.device(), &[0], &[0], &fan, &[1], fan.len() as i32, sampler_kind, &[Some(inputs)], flt)?;
    let (ns, ne, lo) = run(&call, 1, 1, fan.len(), stream_of(col_ptrs))?;
    Ok((call.samples[0].narrow(0, 0, ns[0]), call.rows[0].narrow(0, 0, ne[0]), call.cols[0].narrow(0, 0, ne[0]),
        call.eidx[0].narrow(0, 0, ne[0]), lo.into_iter().next().unwrap_or_default()))
}

/// replaces python.rs:293-394.  Relations are visited in `edge_types` order (the reference iterates a HashMap: quirk Q6);
/// relations absent from `num_neighbors` are inactive and return empty outputs (neighbor_sampling.rs:280-285).
#[allow(clippy::type_complexity)]
pub fn neighbor_sampling_heterogenous_cuda(
    node_types: &[String], edge_types: &[(String, String, String)], col_ptrs: &HashMap<String, Tensor>,
    row_indices: &HashMap<String, Tensor>, inputs: &HashMap<String, Tensor>, num_neighbors: &HashMap<String, Vec<usize>>,
    num_hops: usize, sampler_kind: i32, weights: Option<&HashMap<String, Tensor>>)
    -> TensorResult<(HashMap<String, Tensor>, HashMap<String, Tensor>, HashMap<String, Tensor>, HashMap<String, Tensor>,
                     HashMap<String, Vec<(i64, i64, i64)>>)> {
    let rel = |e: &(String, String, String)| format!("{}__{}__{}", e.0, e.1, e.2);   // neighbor_sampling.rs:255-258
    let tix: HashMap<&str, i32> = node_types.iter().enumerate().map(|(i, t)| (t.as_str(), i as i32)).collect();
    let rels: Vec<String> = edge_types.iter().map(rel).collect();
    let cp: Vec<&Tensor> = rels.iter().map(|r| &col_ptrs[r]).collect();
    let ri: Vec<&Tensor> = rels.iter().map(|r| &row_indices[r]).collect();
    let ws: Option<Vec<&Tensor>> = weights.map(|w| rels.iter().map(|r| &w[r]).collect());
    let graph = graph_for(&cp, &ri, ws.as_deref())?;
    let (src, dst): (Vec<i32>, Vec<i32>) = edge_types.iter().map(|e| (tix[e.0.as_str()], tix[e.2.as_str()])).unzip();
    let active: Vec<u8> = rels.iter().map(|r| num_neighbors.contains_key(r) as u8).collect();
    let mut fan = vec![0i64; rels.len() * num_hops];
    for (i, r) in rels.iter().enumerate() {
        if let Some(ks) = num_neighbors.get(r) { for h in 0..num_hops { fan[i * num_hops + h] = ks[h] as i64; } }  // index panic as :295
    }
    let inp: Vec<Option<&Tensor>> = node_types.iter().map(|t| inputs.get(t)).collect();
    let dev = cp[0].device();
    let call = make_call(graph, dev, &src, &dst, &fan, &active, num_hops as i32, sampler_kind, &inp, None)?;
    let (ns, ne, lo) = run(&call, node_types.len(), rels.len(), num_hops, stream_of(cp[0]))?;
    let samples = node_types.iter().enumerate().map(|(i, t)| (t.clone(), call.samples[i].narrow(0, 0, ns[i]))).collect();
    let pick = |v: &Vec<Tensor>| rels.iter().enumerate().map(|(i, r)| (r.clone(), v[i].narrow(0, 0, ne[i]))).collect();
    let offsets = rels.iter().enumerate().map(|(i, r)| (r.clone(), if active[i] != 0 { lo[i].clone() } else { vec![] })).collect();
    Ok((samples, pick(&call.rows), pick(&call.cols), pick(&call.eidx), offsets))
}

/// replaces python.rs:592-607 (random_walk over CSR; the handle owns the int32 replica of col_indices)
pub fn random_walk_cuda(row_ptrs: &Tensor, col_indices: &Tensor, start: &Tensor, walk_length: i64, p: f32, q: f32)
    -> TensorResult<Tensor> {
    let dev = row_ptrs.device();
    let graph = graph_for(&[row_ptrs], &[col_indices], None)?;
    let walks = Tensor::empty(&[start.size()[0], walk_length + 1], (Kind::Int64, dev));
    let stats = Tensor::empty(&[2], (Kind::Int64, dev));
    ffi::check(unsafe { ffi::tchgeo_random_walk_graph(graph.handle, 0, cuda_ptr::<i64>(start, Kind::Int64)?, start.numel() as i64,
        walk_length, p, q, random::rng_get().next_u64(), 0, walks.data_ptr() as *mut i64, stats.data_ptr() as *mut i64,
        std::ptr::null_mut(), stream_of(row_ptrs)) })?;
    Ok(walks)
}
