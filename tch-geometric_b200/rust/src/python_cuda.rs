//! Bodies that replace the CPU calls in src/python.rs for the hot path (shown for the homogeneous
//! sampler and random_walk; to_csc/to_csr and the heterogeneous sampler follow the same pattern).
//! Signatures, argument extraction and return layouts of python.rs are untouched.
//! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no cargo/rustc).
use tch::{Device, Kind, Tensor};
use crate::cuda_ffi as ffi;
use crate::utils::{random, TensorConversionError, TensorResult};

/// CUDA counterpart of try_tensor_to_slice (src/utils/tensor.rs:50-59): device + dtype check, raw pointer.
fn cuda_ptr<T>(t: &Tensor, kind: Kind) -> TensorResult<*const T> {
    if !matches!(t.device(), Device::Cuda(_)) { return Err(TensorConversionError::InvalidDevice(Device::Cuda(0))); }
    if t.kind() != kind { return Err(TensorConversionError::InvalidDType(kind, t.kind())); }
    Ok(t.data_ptr() as *const T)
}

/// replaces python.rs:202-262 (rng_get, slices, CscGraph::new, algo call, Vec -> Tensor copies)
pub fn neighbor_sampling_homogenous_cuda(col_ptrs: &Tensor, row_indices: &Tensor, inputs: &Tensor,
                                         num_neighbors: &[usize], sampler_kind: i32, weights: Option<&Tensor>)
    -> TensorResult<(Tensor, Tensor, Tensor, Tensor, Vec<(i64, i64, i64)>)> {
    use rand::RngCore;
    let seed = random::rng_get().next_u64();             // the global RNG still forks one child per call
    let dev = col_ptrs.device();
    let (zero, one_b) = (0i32, 1i64);
    let fan: Vec<i64> = num_neighbors.iter().map(|&k| k as i64).collect();
    let (cp, ri, inp) = (cuda_ptr::<i64>(col_ptrs, Kind::Int64)?, cuda_ptr::<i64>(row_indices, Kind::Int64)?,
                         cuda_ptr::<i64>(inputs, Kind::Int64)?);
    let w = match weights { Some(t) => cuda_ptr::<f64>(t, Kind::Double)?, None => std::ptr::null() };
    let (num_cols, s) = (col_ptrs.numel() as i64 - 1, inputs.numel() as i64);
    let mut a: ffi::tchgeo_sampling_args = unsafe { std::mem::zeroed() };
    a.num_node_types = 1; a.num_rels = 1; a.num_hops = fan.len() as i32; a.sampler_kind = sampler_kind;
    a.rel_src = &zero; a.rel_dst = &zero;
    a.col_ptrs = &cp; a.num_cols = &num_cols; a.row_indices = &ri; a.weights = &w;
    a.fanouts = fan.as_ptr(); a.num_batches = one_b; a.inputs = &inp; a.seeds_per_batch = &s; a.seed = seed;
    let (mut cap_n, mut cap_e) = (0i64, 0i64);
    ffi::check(unsafe { ffi::tchgeo_neighbor_sampling_capacity(&a, &mut cap_n, &mut cap_e) })?;
    let samples = Tensor::empty(&[cap_n], (Kind::Int64, dev));
    let (rows, cols, eidx) = (Tensor::empty(&[cap_e], (Kind::Int64, dev)), Tensor::empty(&[cap_e], (Kind::Int64, dev)),
                              Tensor::empty(&[cap_e], (Kind::Int64, dev)));
    let (ps, pr, pc, pe) = (samples.data_ptr() as *mut i64, rows.data_ptr() as *mut i64,
                            cols.data_ptr() as *mut i64, eidx.data_ptr() as *mut i64);
    a.samples = &ps; a.samples_stride = &cap_n; a.rows = &pr; a.cols = &pc; a.edge_index = &pe; a.edges_stride = &cap_e;
    let ws = Tensor::empty(&[unsafe { ffi::tchgeo_neighbor_sampling_workspace_bytes(&a) } as i64], (Kind::Uint8, dev));
    a.workspace = ws.data_ptr(); a.workspace_bytes = ws.numel();
    let (mut ns, mut ne) = (0i64, 0i64);
    let mut lo = vec![0i64; 3 * fan.len().max(1)];
    a.samples_len = &mut ns; a.edges_len = &mut ne; a.layer_offsets = lo.as_mut_ptr();
    a.stream = std::ptr::null_mut();                     // torch's current stream via at::cuda in a real build
    ffi::check(unsafe { ffi::tchgeo_neighbor_sampling(&a) })?;
    let layer_offsets = (0..fan.len()).map(|h| (lo[3 * h], lo[3 * h + 1], lo[3 * h + 2])).collect();
    Ok((samples.narrow(0, 0, ns), rows.narrow(0, 0, ne), cols.narrow(0, 0, ne), eidx.narrow(0, 0, ne), layer_offsets))
}

/// replaces python.rs:592-607
pub fn random_walk_cuda(row_ptrs: &Tensor, col_indices: &Tensor, start: &Tensor, walk_length: i64, p: f32, q: f32)
    -> TensorResult<Tensor> {
    use rand::RngCore;
    let dev = row_ptrs.device();
    let walks = Tensor::empty(&[start.size()[0], walk_length + 1], (Kind::Int64, dev));
    let stats = Tensor::empty(&[2], (Kind::Int64, dev));
    ffi::check(unsafe { ffi::tchgeo_random_walk(
        cuda_ptr::<i64>(row_ptrs, Kind::Int64)?, row_ptrs.numel() as i64 - 1, cuda_ptr::<i64>(col_indices, Kind::Int64)?,
        cuda_ptr::<i64>(start, Kind::Int64)?, start.numel() as i64, walk_length, p, q,
        random::rng_get().next_u64(), 0, walks.data_ptr() as *mut i64, stats.data_ptr() as *mut i64,
        std::ptr::null_mut(), std::ptr::null_mut()) })?;
    Ok(walks)
}
