"""Device-side versions of the data steps that sit directly either side of the sampling loop."""
