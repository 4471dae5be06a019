// Replacement for the trait in src/scoring.rs:11-19: ONE provided method is added; `energy` keeps its signature, so
// DFIRE / DNA / PYDOCK (src/dfire.rs:264, src/dna.rs:410, src/pydock.rs:425) compile unchanged and inherit the default.
pub trait Score {
    fn energy(
        &self,
        translation: &[f64],
        rotation: &Quaternion,
        rec_nmodes: &[f64],
        lig_nmodes: &[f64],
    ) -> f64;

    /// All poses of a GSO step at once.  `poses` is row-major `[n][pose_len]`:
    /// tx, ty, tz, qw, qx, qy, qz, `rec_num_anm` receptor extents, then the ligand extents (src/swarm.rs:33-51).
    /// Default: the reference's behaviour, one `energy` call per row.  `CudaScore` overrides it with one
    /// `ld_score_batch` call (include/lightdock_b200.h).
    fn energy_batch(&self, poses: &[f64], pose_len: usize, rec_num_anm: usize) -> Vec<f64> {
        poses
            .chunks_exact(pose_len)
            .map(|row| {
                let rotation = Quaternion::new(row[3], row[4], row[5], row[6]);
                self.energy(&row[0..3], &rotation, &row[7..7 + rec_num_anm], &row[7 + rec_num_anm..])
            })
            .collect()
    }
}
