"""Host-side mirror of the reference's module tree for the post-processing path: same module paths,
function names, argument meaning and error behaviour as calmiLovesAI/ComputerVision.pytorch's
`core/...`, with the arithmetic routed to libcvpp's CUDA kernels."""
