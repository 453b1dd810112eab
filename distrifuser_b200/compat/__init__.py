"""Stand-ins for the few diffusers==0.24.0 classes the reference's API surface needs, used only when `diffusers`
itself is not importable (this image has neither the package nor a network)."""
