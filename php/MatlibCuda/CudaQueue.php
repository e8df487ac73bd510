<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

/**
 * The command queue NDArrayCuda / LinearAlgebraCuda are constructed with = the library stream of
 * librindow_b200.so.  The driver is stream-ordered: launches are submitted eagerly and complete in
 * order, so finish() is rb200_sync() and getContext() only identifies the device
 * (cf. the OpenCL CommandQueue the reference passes to NDArrayCL, src/NDArrayCL.php:68-90).
 * Mirrors rindow_math_matrix_b200/ndarray_cuda.py (CudaQueue, CudaContext).
 */
class CudaQueue
{
    protected CudaContext $context;
    /** @var array<string,mixed>|null */
    protected ?array $options;

    public function __construct(?array $options=null)
    {
        CudaFFI::lib();
        $this->context = new CudaContext();
        $this->options = $options;
    }

    public function getContext() : CudaContext { return $this->context; }
    public function finish() : void { CudaFFI::check(CudaFFI::lib()->rb200_sync()); }
    public function flush() : void {}
}
