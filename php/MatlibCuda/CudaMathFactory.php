<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

/** `mathFactory` slot of AbstractMatlibService: isAvailable(), Math(). */
class CudaMathFactory
{
    public function isAvailable() : bool { return CudaFFI::isAvailable(); }
    public function Math() : object { return new CudaMath(); }
}
