<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

use Interop\Polite\Math\Matrix\BLAS;
use Interop\Polite\Math\Matrix\NDArray;
use Interop\Polite\Math\Matrix\Buffer;
use InvalidArgumentException;
use FFI;
use Rindow\Math\Matrix\Drivers\MatlibPHP\PhpBlas;
use Rindow\Math\Matrix\Drivers\MatlibPHP\Utils;

/**
 * blas() object of the B200 service.  gemm runs on the device; every other BLAS method is
 * inherited from PhpBlas and therefore runs on the host mirror of the CudaBuffer (ArrayAccess),
 * exactly as SURVEY.md section 8(b) prescribes for the rows marked "next".
 */
class CudaBlas extends PhpBlas
{
    use Utils;

    /** 0 auto (library default: 3xTF32), 1 FP32 SIMT, 2 TF32, 3 3xTF32 */
    protected int $gemmMode = 0;

    public function setGemmMode(int $mode) : void { $this->gemmMode = $mode; }
    public function getConfig() : string { return 'CudaBlas sm_100a'; }
    public function getCorename() : string { return 'B200'; }
    public function hasOmatcopy() : bool { return true; }

    public function gemm(
        int $order, int $transA, int $transB, int $m, int $n, int $k,
        float|object $alpha,
        Buffer $A, int $offsetA, int $ldA,
        Buffer $B, int $offsetB, int $ldB,
        float|object $beta,
        Buffer $C, int $offsetC, int $ldC ) : void
    {
        $dtype = $A->dtype();
        if (!($A instanceof CudaBuffer) || ($dtype != NDArray::float32 && $dtype != NDArray::float64)) {
            parent::gemm($order,$transA,$transB,$m,$n,$k,$alpha,$A,$offsetA,$ldA,$B,$offsetB,$ldB,$beta,$C,$offsetC,$ldC);
            return;
        }
        // same validation, same messages as PhpBlas::gemm (PhpBlas.php:861-881)
        if ($order == BLAS::ColMajor) {
            [$m,$n] = [$n,$m];
        } elseif ($order != BLAS::RowMajor) {
            throw new InvalidArgumentException('Invalid Order type');
        }
        [$tA,$conjA] = $this->codeToTrans($transA);
        [$tB,$conjB] = $this->codeToTrans($transB);
        $this->assertShapeParameter('m',$m);
        $this->assertShapeParameter('n',$n);
        $this->assertShapeParameter('k',$k);
        $this->assertMatrixBufferSpec("A", $A, (!$tA) ? $m : $k, (!$tA) ? $k : $m, $offsetA, $ldA);
        $this->assertMatrixBufferSpec("B", $B, (!$tB) ? $k : $n, (!$tB) ? $n : $k, $offsetB, $ldB);
        $this->assertMatrixBufferSpec("C", $C, $m, $n, $offsetC, $ldC);
        $alpha = $this->cleanFloatNumber($alpha,'alpha');
        $beta = $this->cleanFloatNumber($beta,'beta');
        $ffi = CudaFFI::lib();
        $a = $A->devicePtr(); $b = $B->devicePtr(); $c = $C->devicePtrRW();
        if ($dtype == NDArray::float32) {
            CudaFFI::check($ffi->rb200_sgemm(BLAS::RowMajor,$transA,$transB,$m,$n,$k,$alpha,
                $ffi->cast('float*',$a),$offsetA,$ldA,$ffi->cast('float*',$b),$offsetB,$ldB,
                $beta,$ffi->cast('float*',$c),$offsetC,$ldC,$this->gemmMode));
        } else {
            CudaFFI::check($ffi->rb200_dgemm(BLAS::RowMajor,$transA,$transB,$m,$n,$k,$alpha,
                $ffi->cast('double*',$a),$offsetA,$ldA,$ffi->cast('double*',$b),$offsetB,$ldB,
                $beta,$ffi->cast('double*',$c),$offsetC,$ldC));
        }
    }
}
