/* oracle/ref_stubs: empty stand-in for the Win32 header OR/kernels.cuh:13 includes */
