<?php
namespace RindowTest\Math\Matrix\Drivers\MatlibCuda\CudaBlasTest;

use Rindow\Math\Matrix\Drivers\Service;
use Rindow\Math\Matrix\Drivers\MatlibCuda\MatlibCuda;
use RindowTest\Math\Matrix\Drivers\MatlibPHP\PhpBlasTest\PhpBlasTest as ORGTest;

/**
 * The reference's driver-level BLAS suite (PhpBlasTest.php, 4307 lines: raw (n, buf, off, ld) calls
 * and the buffer-overflow messages of :1988-2155) over CudaBlas -- same override point as
 * PhpBlasPHPModeTest.php:13-18.
 */
class CudaBlasTest extends ORGTest
{
    public function setUp() : void
    {
        $this->service = new MatlibCuda();
        if($this->service->serviceLevel()<Service::LV_ADVANCED) {
            $this->markTestSkipped('librindow_b200.so / B200 not available');
        }
    }

    public function getBlas()
    {
        return $this->service->blas();
    }
}
