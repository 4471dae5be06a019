//! cuda_score.rs — `impl Score` backed by the B200 library (include/lightdock_b200.h).
//! Add `pub mod cuda_score;` to src/lib.rs.  Replaces the bodies of DFIRE::energy (src/dfire.rs:265-362),
//! DNA::energy (src/dna.rs:411-529) and PYDOCK::energy (src/pydock.rs:426-544); the docking models are still
//! built by the existing constructors and handed over as plain arrays.
use super::qt::Quaternion;
use super::scoring::Score;
use std::collections::HashMap;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct LdMoleculeDesc {
    pub n_atoms: i32,
    pub coords: *const f64,
    pub dfire_type: *const i32,
    pub ele_charge: *const f64,
    pub vdw_energy: *const f64,
    pub vdw_radius: *const f64,
    pub n_modes: i32,
    pub modes: *const f64,
    pub n_restraints: i32,
    pub rst_offsets: *const i32,
    pub rst_atoms: *const i32,
    pub n_membrane: i32,
    pub membrane: *const i32,
}

#[repr(C)]
pub struct LdComplexDesc {
    pub method: i32, // 0 DFIRE, 1 DNA, 2 PYDOCK
    pub use_anm: i32,
    pub receptor: LdMoleculeDesc,
    pub ligand: LdMoleculeDesc,
    pub dfire_potential: *const f64,
    pub device: i32,
    pub reserved: i32,
}

#[link(name = "lightdock_b200")]
extern "C" {
    fn ld_create(desc: *const LdComplexDesc, out: *mut *mut c_void) -> c_int;
    fn ld_destroy(h: *mut c_void) -> c_int;
    fn ld_pose_len(h: *const c_void) -> c_int;
    fn ld_score_batch(h: *mut c_void, n_poses: i64, poses: *const f64, energies: *mut f64) -> c_int;
    fn ld_last_error() -> *const c_char;
    // optional: the whole GSO loop on the device (see `DeviceGso` at the end of this file)
    fn ld_gso_create(h: *mut c_void, n_swarms: i32, n_glowworms: i32, positions: *const f64, seeds: *const u64, out: *mut *mut c_void) -> c_int;
    fn ld_gso_run(g: *mut c_void, n_steps: i32) -> c_int;
    fn ld_gso_state(g: *mut c_void, poses: *mut f64, luciferin: *mut f64, vision_range: *mut f64, scoring: *mut f64, n_neighbors: *mut i32, failed_step: *mut i32) -> c_int;
    fn ld_gso_destroy(g: *mut c_void) -> c_int;
}

/// Numeric view of one docking model (what DFIREDockingModel / DNADockingModel already hold).
pub struct ModelArrays<'a> {
    pub coordinates: &'a [[f64; 3]],
    pub dfire_atoms: Option<&'a [usize]>,
    pub ele_charges: Option<&'a [f64]>,
    pub vdw_charges: Option<&'a [f64]>,
    pub vdw_radii: Option<&'a [f64]>,
    pub nmodes: &'a [f64],
    pub num_anm: usize,
    pub active_restraints: &'a HashMap<String, Vec<usize>>,
    pub membrane: &'a [usize],
}

pub struct CudaScore {
    handle: *mut c_void,
    pose_len: usize,
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ld_last_error()).to_string_lossy().into_owned() }
}

impl CudaScore {
    pub fn new(method: i32, rec: &ModelArrays, lig: &ModelArrays, potential: &[f64], use_anm: bool, device: i32) -> Box<dyn Score> {
        struct Owned { types: Vec<i32>, off: Vec<i32>, idx: Vec<i32>, mem: Vec<i32> }
        fn own(m: &ModelArrays) -> Owned {
            let types = m.dfire_atoms.map(|a| a.iter().map(|&t| t as i32).collect()).unwrap_or_default();
            let (mut off, mut idx) = (vec![0i32], Vec::new());
            for atoms in m.active_restraints.values() {
                idx.extend(atoms.iter().map(|&a| a as i32));
                off.push(idx.len() as i32);
            }
            Owned { types, off, idx, mem: m.membrane.iter().map(|&a| a as i32).collect() }
        }
        fn desc(m: &ModelArrays, o: &Owned, use_anm: bool) -> LdMoleculeDesc {
            let p = |s: Option<&[f64]>| s.map_or(std::ptr::null(), |v| v.as_ptr());
            LdMoleculeDesc {
                n_atoms: m.coordinates.len() as i32,
                coords: m.coordinates.as_ptr() as *const f64,
                dfire_type: if o.types.is_empty() { std::ptr::null() } else { o.types.as_ptr() },
                ele_charge: p(m.ele_charges), vdw_energy: p(m.vdw_charges), vdw_radius: p(m.vdw_radii),
                n_modes: if use_anm { m.num_anm as i32 } else { 0 },
                modes: if m.nmodes.is_empty() { std::ptr::null() } else { m.nmodes.as_ptr() },
                n_restraints: (o.off.len() - 1) as i32, rst_offsets: o.off.as_ptr(),
                rst_atoms: if o.idx.is_empty() { std::ptr::null() } else { o.idx.as_ptr() },
                n_membrane: o.mem.len() as i32,
                membrane: if o.mem.is_empty() { std::ptr::null() } else { o.mem.as_ptr() },
            }
        }
        let (ro, lo) = (own(rec), own(lig));
        let d = LdComplexDesc {
            method, use_anm: use_anm as i32, receptor: desc(rec, &ro, use_anm), ligand: desc(lig, &lo, use_anm),
            dfire_potential: if potential.is_empty() { std::ptr::null() } else { potential.as_ptr() },
            device, reserved: 0,
        };
        let mut handle = std::ptr::null_mut();
        if unsafe { ld_create(&d, &mut handle) } != 0 {
            panic!("lightdock_b200: {}", last_error()); // the reference constructors panic on bad input too
        }
        let pose_len = unsafe { ld_pose_len(handle) } as usize;
        Box::new(CudaScore { handle, pose_len })
    }

}

impl Score for CudaScore {
    fn energy(&self, translation: &[f64], rotation: &Quaternion, rec_nmodes: &[f64], lig_nmodes: &[f64]) -> f64 {
        let mut row = Vec::with_capacity(self.pose_len);
        row.extend_from_slice(&translation[..3]);
        row.extend_from_slice(&[rotation.w, rotation.x, rotation.y, rotation.z]);
        row.extend_from_slice(rec_nmodes);
        row.extend_from_slice(lig_nmodes);
        assert_eq!(row.len(), self.pose_len);
        self.energy_batch(&row, self.pose_len, rec_nmodes.len())[0]
    }

    /// Overrides the trait's provided method (scoring_trait.patch.rs): all poses of a step in ONE launch sequence.
    /// Rows of `pose_len` f64 (tx,ty,tz,qw,qx,qy,qz, rec extents, lig extents); the library knows the split.
    fn energy_batch(&self, poses: &[f64], pose_len: usize, _rec_num_anm: usize) -> Vec<f64> {
        assert_eq!(pose_len, self.pose_len, "pose rows do not match the scoring object's pose length");
        let n = poses.len() / self.pose_len;
        let mut out = vec![0.0f64; n];
        if unsafe { ld_score_batch(self.handle, n as i64, poses.as_ptr(), out.as_mut_ptr()) } != 0 {
            panic!("lightdock_b200: {}", last_error());
        }
        out
    }
}

impl Drop for CudaScore {
    fn drop(&mut self) {
        unsafe { ld_destroy(self.handle) };
    }
}


/// Optional: `GSO::run` (src/lib.rs:46-58) with the whole step on the device.  One swarm per `positions` chunk of
/// `n_glowworms * pose_len` values, each seeded like `StdRng::seed_from_u64(seed)`.  `run(steps)` advances every swarm;
/// `state()` returns what `Swarm::save` prints.  Trajectories equal the host loop's to the last bit or two (CUDA's
/// acos/sin inside slerp), so upstream would keep `GSO::run` as the default and offer this behind a flag.
pub struct DeviceGso {
    gso: *mut c_void,
    n: usize,
    pose_len: usize,
}

impl DeviceGso {
    pub fn new(score: &CudaScore, n_swarms: usize, n_glowworms: usize, positions: &[f64], seeds: &[u64]) -> DeviceGso {
        assert_eq!(positions.len(), n_swarms * n_glowworms * score.pose_len);
        assert_eq!(seeds.len(), n_swarms);
        let mut gso: *mut c_void = std::ptr::null_mut();
        let rc = unsafe { ld_gso_create(score.handle, n_swarms as i32, n_glowworms as i32, positions.as_ptr(), seeds.as_ptr(), &mut gso) };
        if rc != 0 {
            panic!("ld_gso_create: {}", last_error());
        }
        DeviceGso { gso, n: n_swarms * n_glowworms, pose_len: score.pose_len }
    }

    pub fn run(&mut self, steps: u32) {
        if unsafe { ld_gso_run(self.gso, steps as i32) } != 0 {
            panic!("ld_gso_run: {}", last_error());
        }
    }

    /// (poses, luciferin, vision_range, scoring, n_neighbors)
    pub fn state(&self) -> (Vec<f64>, Vec<f64>, Vec<f64>, Vec<f64>, Vec<i32>) {
        let mut poses = vec![0.0; self.n * self.pose_len];
        let (mut lum, mut vis, mut sc) = (vec![0.0; self.n], vec![0.0; self.n], vec![0.0; self.n]);
        let mut nn = vec![0i32; self.n];
        let rc = unsafe {
            ld_gso_state(self.gso, poses.as_mut_ptr(), lum.as_mut_ptr(), vis.as_mut_ptr(), sc.as_mut_ptr(), nn.as_mut_ptr(), std::ptr::null_mut())
        };
        if rc != 0 {
            panic!("ld_gso_state: {}", last_error());
        }
        (poses, lum, vis, sc, nn)
    }
}

impl Drop for DeviceGso {
    fn drop(&mut self) {
        unsafe { ld_gso_destroy(self.gso) };
    }
}
