"""Mirror of the single hot-path entry of memflow.unfolding_flow: Compute_ParticlesTensor.get_PS."""
