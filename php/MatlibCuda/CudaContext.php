<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

/**
 * What CudaQueue::getContext() returns: identifies the device the library stream belongs to (the stand-in for the
 * OpenCL Context NDArrayCL keeps, src/NDArrayCL.php:54,86).  Mirrors rindow_math_matrix_b200/ndarray_cuda.py (CudaContext).
 */
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
