<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

/** `bufferFactory` of AbstractMatlibService (src/Drivers/AbstractMatlibService.php:43-75). */
class CudaBufferFactory
{
    public function isAvailable() : bool { return CudaFFI::isAvailable(); }
    public function Buffer(int $size, int $dtype) : object { return new CudaBuffer($size, $dtype); }
}
