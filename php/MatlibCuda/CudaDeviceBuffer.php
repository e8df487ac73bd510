<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

use Interop\Polite\Math\Matrix\DeviceBuffer;
use Interop\Polite\Math\Matrix\LinearBuffer;
use Interop\Polite\Math\Matrix\NDArray;
use Rindow\Math\Matrix\Complex;
use FFI;
use InvalidArgumentException;
use LogicException;
use RuntimeException;

/**
 * DeviceBuffer of the B200 driver (contract used by src/NDArrayCL.php:147-168, 261-270, 416-441, 503-512): bytes(),
 * dtype(), value_size(), queue-ordered read / write / copy / fill with BYTE sizes and offsets.  HBM only -- there is
 * no pinned mirror; every bulk transfer is explicit.  Single elements can still be peeked / poked (`$buf[$i]` costs one
 * small d2h / h2d), which keeps LinearAlgebra's four host reads (`$XX[$offX+$i]`, src/LinearAlgebra.php:488-511,
 * 1645-1658) and the driver's host-read parameters working on device arrays.
 * Mirrors rindow_math_matrix_b200/ndarray_cuda.py (CudaDeviceBuffer).
 */
class CudaDeviceBuffer extends CudaBuffer implements DeviceBuffer
{
    protected int $flags;

    public function __construct(
        object $context, int $bytes, int $flags=0, ?object $hostBuffer=null, int $hostOffset=0, ?int $dtype=null)
    {
        if($dtype===null) {
            if($hostBuffer===null) { throw new InvalidArgumentException('dtype is required'); }
            $dtype = $hostBuffer->dtype();
        }
        $t = self::table();
        if(!isset($t[$dtype])) { throw new InvalidArgumentException('Invalid data type'); }
        $vs = $t[$dtype][1];
        if($bytes % $vs) { throw new InvalidArgumentException('bytes is not a multiple of the value size'); }
        parent::__construct(intdiv($bytes, $vs), $dtype, zero:false);
        $this->flags = $flags;
        if($hostBuffer!==null) {                     // CL_MEM_COPY_HOST_PTR: starts as a copy of hostBuffer[hostOffset:]
            $this->write(null, $hostBuffer, $bytes, 0, $hostOffset*$vs);
        }
    }

    public function bytes() : int { return $this->size * $this->valueSize; }
    public function flags() : int { return $this->flags; }

    protected function mirror() { throw new LogicException('CudaDeviceBuffer has no host mirror: use read()'); }
    public function ensureMirror() : void { throw new LogicException('CudaDeviceBuffer has no host mirror'); }

    public function offsetGet($offset) : mixed
    {
        if(!$this->offsetExists($offset)) { throw new RuntimeException('Index invalid or out of range'); }
        $ffi = CudaFFI::lib();
        if($this->isComplexDtype()) {                // interleaved real / imag pair of the component type
            $pair = $ffi->new($this->ctype.'[2]');
            CudaFFI::check($ffi->rb200_buffer_d2h($pair, $this->dptr, $offset*$this->valueSize, $this->valueSize));
            return new Complex((float)$pair[0], (float)$pair[1]);
        }
        $one = $ffi->new($this->ctype);
        CudaFFI::check($ffi->rb200_buffer_d2h(FFI::addr($one), $this->dptr, $offset*$this->valueSize, $this->valueSize));
        return ($this->dtype == NDArray::bool) ? ($one->cdata ? true : false) : $one->cdata;
    }

    /** one element of this dtype on the host, ready to be passed as `const void*` */
    protected function hostElement(mixed $value) : object
    {
        $ffi = CudaFFI::lib();
        if($this->isComplexDtype()) {
            if(!is_object($value) || !property_exists($value, 'real') || !property_exists($value, 'imag')) {
                throw new InvalidArgumentException('Cannot convert to complex number.: '.gettype($value));
            }
            $pair = $ffi->new($this->ctype.'[2]');
            $pair[0] = $value->real;
            $pair[1] = $value->imag;
            return $pair;
        }
        $one = $ffi->new($this->ctype.'[1]');
        $one[0] = $value;
        return $one;
    }

    public function offsetSet($offset, $value) : void
    {
        if(!$this->offsetExists($offset)) { throw new RuntimeException('Index invalid or out of range'); }
        CudaFFI::check(CudaFFI::lib()->rb200_buffer_h2d($this->dptr, $offset*$this->valueSize,
            $this->hostElement($value), $this->valueSize));
        $this->devDirty = true;
    }

    /** device -> host buffer (src/NDArrayCL.php:261-270).  A CudaBuffer partner receives the bytes on the device side. */
    public function read(?object $queue, object $hostBuffer, int $size=0, int $offset=0, int $hostoffset=0,
        ?bool $blocking_read=null, ?object $events=null, ?object $waitEvents=null) : void
    {
        if($waitEvents!==null) { $waitEvents->wait(); }
        $size = $size ?: ($this->bytes() - $offset);
        $ffi = CudaFFI::lib();
        if($hostBuffer instanceof CudaBuffer) {
            CudaFFI::check($ffi->rb200_buffer_d2d($hostBuffer->devicePtrRW(), $hostoffset, $this->dptr, $offset, $size));
            if($blocking_read===null || $blocking_read) { CudaFFI::check($ffi->rb200_sync()); }
        } elseif($hostBuffer instanceof LinearBuffer && method_exists($hostBuffer, 'addr')) {
            // an FFI-backed host buffer of another driver (rindow-math-buffer-ffi): raw address of element 0
            CudaFFI::check($ffi->rb200_buffer_d2h($hostBuffer->addr(0), $this->dptr, $offset, $size));
        } else {
            $tmp = $ffi->new('char['.$size.']');
            CudaFFI::check($ffi->rb200_buffer_d2h($tmp, $this->dptr, $offset, $size));
            $hostBuffer->load(FFI::string($tmp, $size));       // PhpBuffer::load, hostoffset 0 only
        }
        if($events!==null) { $events->record(); }
    }

    public function write(?object $queue, object $hostBuffer, int $size=0, int $offset=0, int $hostoffset=0,
        ?bool $blocking_write=null, ?object $events=null, ?object $waitEvents=null) : void
    {
        if($waitEvents!==null) { $waitEvents->wait(); }
        $ffi = CudaFFI::lib();
        if($hostBuffer instanceof CudaBuffer) {
            $size = $size ?: min(count($hostBuffer)*$hostBuffer->value_size() - $hostoffset, $this->bytes() - $offset);
            CudaFFI::check($ffi->rb200_buffer_d2d($this->dptr, $offset, $hostBuffer->devicePtr(), $hostoffset, $size));
        } else {
            $bytes = $hostBuffer->dump();
            $size = $size ?: min(strlen($bytes) - $hostoffset, $this->bytes() - $offset);
            $tmp = $ffi->new('char['.max(1,$size).']');
            FFI::memcpy($tmp, substr($bytes, $hostoffset, $size), $size);
            CudaFFI::check($ffi->rb200_buffer_h2d($this->dptr, $offset, $tmp, $size));
        }
        $this->devDirty = true;
        if($events!==null) { $events->record(); }
    }

    /** DeviceBuffer::copy(queue, src, size, srcOffset, dstOffset, events, waitEvents) (src/NDArrayCL.php:416-441) */
    public function copy(?object $queue, object $src, int $size=0, int $src_offset=0, int $dst_offset=0,
        ?object $events=null, ?object $waitEvents=null) : void
    {
        if(!($src instanceof CudaBuffer)) {
            throw new InvalidArgumentException('copy source must be a buffer of this driver: '.get_class($src));
        }
        if($waitEvents!==null) { $waitEvents->wait(); }
        $size = $size ?: (count($src)*$src->value_size() - $src_offset);
        CudaFFI::check(CudaFFI::lib()->rb200_buffer_d2d($this->dptr, $dst_offset, $src->devicePtr(), $src_offset, $size));
        $this->devDirty = true;
        if($events!==null) { $events->record(); }
    }

    /** fill $size bytes at byte $offset with the first element of $pattern (a host buffer of this dtype) */
    public function fill(?object $queue, object $pattern, int $size=0, int $offset=0, int $pattern_size=0,
        int $pattern_offset=0, ?object $events=null, ?object $waitEvents=null) : void
    {
        if($waitEvents!==null) { $waitEvents->wait(); }
        $size = $size ?: ($this->bytes() - $offset);
        CudaFFI::check(CudaFFI::lib()->rb200_buffer_fill($this->abiDtype, $this->dptr, intdiv($offset,$this->valueSize),
            intdiv($size,$this->valueSize), $this->hostElement($pattern[$pattern_offset])));
        $this->devDirty = true;
        if($events!==null) { $events->record(); }
    }

    public function dump() : string
    {
        $n = $this->bytes();
        $ffi = CudaFFI::lib();
        $tmp = $ffi->new('char['.max(1,$n).']');
        CudaFFI::check($ffi->rb200_buffer_d2h($tmp, $this->dptr, 0, $n));
        return FFI::string($tmp, $n);
    }

    public function load(string $string) : void
    {
        if(strlen($string) > $this->bytes()) { throw new RuntimeException('Unpack error'); }
        $ffi = CudaFFI::lib();
        $tmp = $ffi->new('char['.max(1,strlen($string)).']');
        FFI::memcpy($tmp, $string, strlen($string));
        CudaFFI::check($ffi->rb200_buffer_h2d($this->dptr, 0, $tmp, strlen($string)));
        $this->devDirty = true;
    }
}
