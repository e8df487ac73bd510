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

class CudaContext
{
    /** @return array<string,int> */
    public function getInfo(mixed $what=null) : array
    {
        $ffi = CudaFFI::lib();
        $dev = $ffi->new('int'); $sm = $ffi->new('int'); $maj = $ffi->new('int'); $mnr = $ffi->new('int');
        $mem = $ffi->new('int64_t');
        CudaFFI::check($ffi->rb200_device_info(\FFI::addr($dev), \FFI::addr($sm), \FFI::addr($maj), \FFI::addr($mnr),
            \FFI::addr($mem)));
        return ['device'=>$dev->cdata, 'sm_count'=>$sm->cdata, 'cc_major'=>$maj->cdata, 'cc_minor'=>$mnr->cdata,
                'total_mem'=>$mem->cdata];
    }
}
