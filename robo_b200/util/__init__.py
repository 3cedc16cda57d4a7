"""Host-side helpers of the B200 path (training-data scaling, the ensemble sampler used by GP-MCMC)."""
