// Replacement body for Swarm::update_luciferin (src/swarm.rs:66-70).  `Score` gains the provided method
//     fn energy_batch(&self, poses: &[f64], pose_len: usize, rec_num_anm: usize) -> Vec<f64>
// (scoring_trait.patch.rs; default: loop over energy()), which CudaScore overrides with one ld_score_batch call.
// Every Glowworm field used here is `pub` upstream (src/glowworm.rs:6-26).  Everything else in swarm.rs /
// glowworm.rs / lib.rs, including the StdRng stream and the order of the `moved` bookkeeping, is untouched.
pub fn update_luciferin(&mut self) {
    if self.glowworms.is_empty() {
        return;
    }
    let scoring = self.glowworms[0].scoring_function;
    let mut rows: Vec<f64> = Vec::new();
    let mut who: Vec<usize> = Vec::new();
    for (i, g) in self.glowworms.iter().enumerate() {
        if g.moved || g.step == 0 {               // src/glowworm.rs:62
            rows.extend_from_slice(&g.translation);
            rows.extend_from_slice(&[g.rotation.w, g.rotation.x, g.rotation.y, g.rotation.z]);
            rows.extend_from_slice(&g.rec_nmodes);
            rows.extend_from_slice(&g.lig_nmodes);
            who.push(i);
        }
    }
    let pose_len = if who.is_empty() { 7 } else { rows.len() / who.len() };
    let rec_num_anm = self.glowworms[0].rec_nmodes.len();
    let scores = scoring.energy_batch(&rows, pose_len, rec_num_anm);
    for (k, &i) in who.iter().enumerate() {
        self.glowworms[i].scoring = scores[k];
    }
    for g in self.glowworms.iter_mut() {          // src/glowworm.rs:70-71, unchanged arithmetic
        g.luciferin = (1.0 - g.rho) * g.luciferin + g.gamma * g.scoring;
        g.step += 1;
    }
}
