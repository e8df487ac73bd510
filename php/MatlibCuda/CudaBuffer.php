<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

use Interop\Polite\Math\Matrix\LinearBuffer;
use Interop\Polite\Math\Matrix\NDArray;
use Rindow\Math\Matrix\Complex;
use FFI;
use InvalidArgumentException;
use RuntimeException;

/**
 * Host-indexable, device-resident Buffer (contract: src/Drivers/MatlibPHP/PhpBuffer.php:19-190).
 * Storage of record is HBM; a pinned host mirror is created on the first ArrayAccess; two dirty
 * flags keep them coherent (flush before a kernel, fetch before a host read).  Mirrors
 * rindow_math_matrix_b200/cuda_buffer.py line for line.
 */
class CudaBuffer implements LinearBuffer
{
    /** NDArray dtype => [C ABI dtype code (include/rindow_b200.h), bytes, FFI element type] */
    protected static function table() : array
    {
        return [
            NDArray::bool    => [0, 1, 'uint8_t'],  NDArray::int8    => [1, 1, 'int8_t'],
            NDArray::int16   => [2, 2, 'int16_t'],  NDArray::int32   => [3, 4, 'int32_t'],
            NDArray::int64   => [4, 8, 'int64_t'],  NDArray::uint8   => [5, 1, 'uint8_t'],
            NDArray::uint16  => [6, 2, 'uint16_t'], NDArray::uint32  => [7, 4, 'uint32_t'],
            NDArray::uint64  => [8, 8, 'uint64_t'], NDArray::float32 => [12, 4, 'float'],
            NDArray::float64 => [13, 8, 'double'],
            // complex storage (PhpBuffer.php:38-41, 60-61): interleaved real / imag pairs; the FFI view is the
            // component type, offsetGet / offsetSet translate between pairs and objects with real / imag
            NDArray::complex64  => [16, 8,  'float'],
            NDArray::complex128 => [17, 16, 'double'],
        ];
    }

    protected int $size;
    protected int $dtype;
    protected int $abiDtype;
    protected int $valueSize;
    protected string $ctype;
    protected $dptr;                 // void* (device)
    protected $hptr = null;          // void* (pinned host mirror)
    protected $host = null;          // typed FFI view of the mirror
    protected bool $hostDirty = false;
    protected bool $devDirty = false;
    protected $uploadEvent = null;   // recorded behind the asynchronous h2d of the mirror (devicePtr)
    protected bool $uploadPending = false;

    public function __construct(int $size, int $dtype, bool $zero=true)
    {
        $t = self::table();
        if (!isset($t[$dtype])) {
            throw new InvalidArgumentException("Invalid data type");
        }
        [$this->abiDtype, $this->valueSize, $this->ctype] = $t[$dtype];
        $this->size = $size;
        $this->dtype = $dtype;
        $ffi = CudaFFI::lib();
        $p = $ffi->new('void*');
        CudaFFI::check($ffi->rb200_buffer_alloc($size * $this->valueSize, FFI::addr($p)));
        $this->dptr = $p;
        if ($zero && $size > 0) {
            // one zero element on the host (complex: a real / imag pair of the component type)
            $z = $ffi->new($this->ctype.'[2]');
            CudaFFI::check($ffi->rb200_buffer_fill($this->abiDtype, $this->dptr, 0, $size, $z));
        }
        $this->devDirty = true;
    }

    public function __destruct()
    {
        $ffi = CudaFFI::lib();
        if ($this->uploadEvent !== null) {
            if ($this->uploadPending) { $ffi->rb200_event_synchronize($this->uploadEvent); }
            $ffi->rb200_event_destroy($this->uploadEvent);
        }
        $ffi->rb200_buffer_free($this->dptr);
        if ($this->hptr !== null) {
            $ffi->rb200_host_free($this->hptr);
        }
    }

    public function dtype() : int { return $this->dtype; }
    public function abiDtype() : int { return $this->abiDtype; }
    /** polite-math dtype constant -> C ABI dtype code (the one mapping table, see table()) */
    public static function abiDtypeOf(int $dtype) : int
    {
        $t = self::table();
        if (!isset($t[$dtype])) {
            throw new InvalidArgumentException('dtype must be type of integer or float: '.$dtype);
        }
        return $t[$dtype][0];
    }
    public function value_size() : int { return $this->valueSize; }
    public function count() : int { return $this->size; }

    /**
     * devicePtr() queues the h2d of the pinned mirror asynchronously, possibly behind long kernels: the host may
     * not modify the mirror before that copy has read it ($X[0]=a; op($X); $X[0]=b must show `a` to op).
     */
    protected function waitUpload() : void
    {
        if ($this->uploadPending) {
            CudaFFI::check(CudaFFI::lib()->rb200_event_synchronize($this->uploadEvent));
            $this->uploadPending = false;
        }
    }

    protected function mirror()
    {
        $ffi = CudaFFI::lib();
        $this->waitUpload();
        if ($this->host === null) {
            $p = $ffi->new('void*');
            CudaFFI::check($ffi->rb200_host_alloc(max(16, $this->size * $this->valueSize), FFI::addr($p)));
            $this->hptr = $p;
            $this->host = $ffi->cast($this->ctype.'['.max(1, $this->size * ($this->isComplexDtype() ? 2 : 1)).']', $p);
            $this->devDirty = true;
        }
        if ($this->devDirty) {
            CudaFFI::check($ffi->rb200_buffer_d2h($this->hptr, $this->dptr, 0, $this->size * $this->valueSize));
            $this->devDirty = false;
        }
        return $this->host;
    }

    /** device pointer for a kernel that reads the buffer */
    public function devicePtr()
    {
        if ($this->hostDirty) {
            $ffi = CudaFFI::lib();
            CudaFFI::check($ffi->rb200_buffer_h2d_async(
                $this->dptr, 0, $this->hptr, $this->size * $this->valueSize));
            if ($this->uploadEvent === null) {
                $ev = $ffi->new('void*');
                CudaFFI::check($ffi->rb200_event_create(FFI::addr($ev)));
                $this->uploadEvent = $ev;
            }
            CudaFFI::check($ffi->rb200_event_record($this->uploadEvent));
            $this->uploadPending = true;
            $this->hostDirty = false;
        }
        return $this->dptr;
    }

    // ---- state the streamed (host-operand) gemm needs: it moves the data itself (CudaBlas::gemmStreamed) ----
    public function hostPending() : bool { return $this->hostDirty; }
    public function hasMirror() : bool { return $this->host !== null; }
    public function mirrorCurrent() : bool { return $this->host !== null && !$this->devDirty; }
    /** [device pointer, pinned host pointer|null] without any implied transfer */
    public function rawPointers() : array { return [$this->dptr, $this->hptr]; }
    public function streamedUploaded() : void { $this->hostDirty = false; }
    public function streamedResult(bool $downloaded) : void { $this->devDirty = !$downloaded; }
    public function ensureMirror() : void
    {
        if ($this->host === null) {
            $ffi = CudaFFI::lib();
            $p = $ffi->new('void*');
            CudaFFI::check($ffi->rb200_host_alloc(max(16, $this->size * $this->valueSize), FFI::addr($p)));
            $this->hptr = $p;
            $this->host = $ffi->cast($this->ctype.'['.max(1, $this->size * ($this->isComplexDtype() ? 2 : 1)).']', $p);
            $this->devDirty = true;
        }
    }

    /** device pointer for a kernel that writes the buffer */
    public function devicePtrRW()
    {
        $p = $this->devicePtr();
        $this->devDirty = true;
        return $p;
    }

    public function offsetExists($offset) : bool { return is_int($offset) && $offset >= 0 && $offset < $this->size; }

    public function offsetGet($offset) : mixed
    {
        if (!$this->offsetExists($offset)) {
            throw new RuntimeException("Index invalid or out of range");
        }
        if ($this->isComplexDtype()) {
            $m = $this->mirror();
            return new Complex((float)$m[2*$offset], (float)$m[2*$offset+1]);      // what C() builds (src/C.php:6-13)
        }
        $v = $this->mirror()[$offset];
        return ($this->dtype == NDArray::bool) ? ($v ? true : false) : $v;
    }

    public function offsetSet($offset, $value) : void
    {
        if (!$this->offsetExists($offset)) {
            throw new RuntimeException("Index invalid or out of range");
        }
        if ($this->isComplexDtype()) {
            if (!is_object($value) || !property_exists($value, 'real') || !property_exists($value, 'imag')) {
                throw new InvalidArgumentException("Cannot convert to complex number.: ".gettype($value));
            }
            $m = $this->mirror();
            $m[2*$offset] = $value->real;
            $m[2*$offset+1] = $value->imag;
        } else {
            $this->mirror()[$offset] = $value;
        }
        $this->hostDirty = true;
    }

    protected function isComplexDtype() : bool
    {
        return $this->dtype == NDArray::complex64 || $this->dtype == NDArray::complex128;
    }

    public function offsetUnset($offset) : void { throw new \LogicException("Illegal Operation"); }

    public function dump() : string
    {
        $this->mirror();
        return FFI::string($this->hptr, $this->size * $this->valueSize);
    }

    public function load(string $string) : void
    {
        if (strlen($string) > $this->size * $this->valueSize) {
            throw new RuntimeException('Unpack error');
        }
        $this->mirror();
        FFI::memcpy($this->hptr, $string, strlen($string));
        $this->hostDirty = true;
    }
}
