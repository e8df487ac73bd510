<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

use Interop\Polite\Math\Matrix\NDArray;
use Interop\Polite\Math\Matrix\Buffer;
use InvalidArgumentException;
use Rindow\Math\Matrix\Drivers\MatlibPHP\PhpMath;

/**
 * math() object of the B200 service.  The hot-path methods run on the device through the C ABI;
 * everything else is inherited from PhpMath and runs on the host mirror of the CudaBuffer.
 */
class CudaMath extends PhpMath
{
    protected function onDevice(Buffer ...$buffers) : bool
    {
        foreach ($buffers as $b) {
            if (!($b instanceof CudaBuffer)) { return false; }
        }
        $dt = $buffers[0]->dtype();
        return $dt == NDArray::float32 || $dt == NDArray::float64;
    }

    public function add(bool $trans,int $m,int $n,float $alpha,
        Buffer $X,int $offsetX,int $incX,Buffer $A,int $offsetA,int $ldA) : void
    {
        if (!$this->onDevice($X,$A)) { parent::add($trans,$m,$n,$alpha,$X,$offsetX,$incX,$A,$offsetA,$ldA); return; }
        $cols = (!$trans) ? $n : $m;
        if ($offsetX+($cols-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        if ($offsetA+($m-1)*$ldA+($n-1)>=count($A))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        CudaFFI::check(CudaFFI::lib()->rb200_add($A->abiDtype(),$trans?1:0,$m,$n,$alpha,
            $X->devicePtr(),$offsetX,$incX,$A->devicePtrRW(),$offsetA,$ldA));
    }

    public function multiply(bool $trans,int $m,int $n,
        Buffer $X,int $offsetX,int $incX,Buffer $A,int $offsetA,int $ldA) : void
    {
        if (!$this->onDevice($X,$A)) { parent::multiply($trans,$m,$n,$X,$offsetX,$incX,$A,$offsetA,$ldA); return; }
        $cols = (!$trans) ? $n : $m;
        if ($offsetX+($cols-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        if ($offsetA+($m-1)*$ldA+($n-1)>=count($A))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        CudaFFI::check(CudaFFI::lib()->rb200_multiply($A->abiDtype(),$trans?1:0,$m,$n,
            $X->devicePtr(),$offsetX,$incX,$A->devicePtrRW(),$offsetA,$ldA));
    }

    protected function checkReduce(int $m,int $n,int $k,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if ($offsetA+$m*$n*$k>count($A))
            throw new InvalidArgumentException('Matrix A specification too large for buffer.');
        if ($offsetB+$m*$k>count($B))
            throw new InvalidArgumentException('Matrix B specification too large for buffer.');
    }

    public function reduceSum(int $m,int $n,int $k,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if (!$this->onDevice($A,$B)) { parent::reduceSum($m,$n,$k,$A,$offsetA,$B,$offsetB); return; }
        $this->checkReduce($m,$n,$k,$A,$offsetA,$B,$offsetB);
        CudaFFI::check(CudaFFI::lib()->rb200_reduce_sum($A->abiDtype(),$m,$n,$k,
            $A->devicePtr(),$offsetA,$B->devicePtrRW(),$offsetB));
    }

    public function reduceMax(int $m,int $n,int $k,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if (!$this->onDevice($A,$B)) { parent::reduceMax($m,$n,$k,$A,$offsetA,$B,$offsetB); return; }
        $this->checkReduce($m,$n,$k,$A,$offsetA,$B,$offsetB);
        CudaFFI::check(CudaFFI::lib()->rb200_reduce_max($A->abiDtype(),$m,$n,$k,
            $A->devicePtr(),$offsetA,$B->devicePtrRW(),$offsetB));
    }

    public function reduceArgMax(int $m,int $n,int $k,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if (!$this->onDevice($A) || !($B instanceof CudaBuffer)) { parent::reduceArgMax($m,$n,$k,$A,$offsetA,$B,$offsetB); return; }
        $this->checkReduce($m,$n,$k,$A,$offsetA,$B,$offsetB);
        CudaFFI::check(CudaFFI::lib()->rb200_reduce_argmax($A->abiDtype(),$m,$n,$k,
            $A->devicePtr(),$offsetA,$B->abiDtype(),$B->devicePtrRW(),$offsetB));
    }

    public function softmax(int $m,int $n,Buffer $A,int $offsetA,int $ldA) : void
    {
        if (!$this->onDevice($A)) { parent::softmax($m,$n,$A,$offsetA,$ldA); return; }
        if ($offsetA+($m-1)*$ldA+($n-1)>=count($A))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        CudaFFI::check(CudaFFI::lib()->rb200_softmax($A->abiDtype(),$m,$n,$A->devicePtrRW(),$offsetA,$ldA));
    }

    public function im2col2d(bool $reverse,Buffer $images,int $images_offset,int $images_size,int $batches,
        int $im_h,int $im_w,int $channels,int $filter_h,int $filter_w,int $stride_h,int $stride_w,
        bool $padding,bool $channels_first,int $dilation_h,int $dilation_w,bool $cols_channels_first,
        Buffer $cols,int $cols_offset,int $cols_size) : void
    {
        if (!($images instanceof CudaBuffer) || !($cols instanceof CudaBuffer)) {
            parent::im2col2d($reverse,$images,$images_offset,$images_size,$batches,$im_h,$im_w,$channels,
                $filter_h,$filter_w,$stride_h,$stride_w,$padding,$channels_first,$dilation_h,$dilation_w,
                $cols_channels_first,$cols,$cols_offset,$cols_size);
            return;
        }
        // same checks and messages as PhpMath::im2col2d (PhpMath.php:2208-2246)
        $images_buf_size = $batches*$im_h*$im_w*$channels;
        if ($images_size!=$images_buf_size || count($images)-$images_offset<$images_buf_size) {
            throw new InvalidArgumentException('images buffer size is invalid');
        }
        $out_h = intdiv(($im_h-($filter_h-1)*$dilation_h-1),$stride_h)+1;
        $out_w = intdiv(($im_w-($filter_w-1)*$dilation_w-1),$stride_w)+1;
        if ($out_h<=0 || $out_w<=0) {
            throw new InvalidArgumentException('Invalid shape or parameters.');
        }
        $out_buf_size = $padding ? $batches*$im_h*$filter_h*$im_w*$filter_w*$channels
                                 : $batches*$out_h*$filter_h*$out_w*$filter_w*$channels;
        if ($cols_size!=$out_buf_size || count($cols)-$cols_offset!=$out_buf_size) {
            throw new InvalidArgumentException('output buffer size is invalid');
        }
        $im = $reverse ? $images->devicePtrRW() : $images->devicePtr();
        $co = $reverse ? $cols->devicePtr() : $cols->devicePtrRW();
        CudaFFI::check(CudaFFI::lib()->rb200_im2col2d($images->abiDtype(),$reverse?1:0,$im,$images_offset,$images_size,
            $batches,$im_h,$im_w,$channels,$filter_h,$filter_w,$stride_h,$stride_w,$padding?1:0,$channels_first?1:0,
            $dilation_h,$dilation_w,$cols_channels_first?1:0,$co,$cols_offset,$cols_size));
    }
}
