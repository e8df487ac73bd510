<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

use Rindow\Math\Matrix\Drivers\MatlibPHP\PhpLapack;

/** `openblasFactory` slot of AbstractMatlibService: isAvailable(), Blas(), Lapack(). */
class CudaBlasFactory
{
    public function isAvailable() : bool { return CudaFFI::isAvailable(); }
    public function Blas() : object { return new CudaBlas(); }
    public function Lapack() : object { return new PhpLapack(); }   // LAPACK stays on the PHP driver
}
