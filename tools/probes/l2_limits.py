"""Query the L2 persistence limits of the device (for the scratch-residency experiment of DESIGN.md section 4)."""
import torch
from cuda import cudart
p = torch.cuda.get_device_properties(0)
print("name", p.name, "L2", p.L2_cache_size, "SMs", p.multi_processor_count)
for name in ("cudaDevAttrMaxPersistingL2CacheSize", "cudaDevAttrMaxAccessPolicyWindowSize", "cudaDevAttrL2CacheSize"):
    err, v = cudart.cudaDeviceGetAttribute(getattr(cudart.cudaDeviceAttr, name), 0)
    print(name, err, v)
err, v = cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitPersistingL2CacheSize)
print("cudaLimitPersistingL2CacheSize (current)", err, v)
